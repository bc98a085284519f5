"""Place the UNMODIFIED reference under baseline/_ref/ so that it travels to the GPU box.

microsoft/esvit has no setup.py / pyproject, so `pip install --target baseline/_ref
/root/reference` has nothing to build; "installing" it is a verbatim copy of its Python sources and experiment YAMLs.
baseline/_ref/ is git-ignored (it never enters the history - the repo holds no reference source) but not
gpurun-ignored, so bench.py's reference arms can import the reference's own modules on the box:
`bench.py --impl reference --device cuda` and the `gpu_reference` entry of the default bench line.

Run by __graft_entry__.build() when /root/reference is present; a no-op elsewhere."""
from __future__ import annotations

import os
import shutil
import sys

SRC = os.environ.get("ESVIT_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
KEEP_EXT = (".py", ".yaml", ".yml", ".txt", ".md")


def install(force: bool = False) -> str:
    if not os.path.isfile(os.path.join(SRC, "main_esvit.py")):
        return DST if os.path.isdir(DST) else ""
    stamp = os.path.join(DST, ".installed")
    if os.path.isfile(stamp) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d not in (".git", "__pycache__", "plot")]
        rel = os.path.relpath(root, SRC)
        for f in files:
            if f.endswith(KEEP_EXT):
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                shutil.copy2(os.path.join(root, f), os.path.join(DST, rel, f))
    open(stamp, "w").write("copied verbatim from %s\n" % SRC)
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
