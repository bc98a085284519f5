// ws = 7 (shifted-)window attention core, FORWARD, on tcgen05 tensor cores with TMEM accumulators.
//
// One persistent CTA per SM serves ONE head and walks over PAIRS of windows: the 2 x 64 slots of a pair are the 128 rows
// of one UMMA tile, so every TMEM lane is one query row and the softmax needs no shuffles at all.
//
//   gather warps (4)  cp.async the q / k / v rows of the pair's slots (pad / roll / partition folded into the addressing,
//                     padded slots = the qkv bias, slots >= 49 zero) into 128B-swizzled tiles [Q] and [K | V] (K and V of a
//                     slot share one 128-byte row), 4 stages with 3 pairs of copies in flight per thread, and an L2 prefetch
//                     four pairs ahead
//   MMA warp (1 thr)  S = Q K^T     tcgen05.mma M=128 N=128 K=32  -> TMEM   (both windows at once; only the two 64 x 64
//                                                                           diagonal blocks are read back)
//                     O = P [K|V]   tcgen05.mma M=128 N=64  K=128 -> TMEM   (P from smem; the [K|V] tile read MN-major as it
//                                                                           lies: columns 32..63 of the result are P V)
//   row warps (2 x 4) thread = query row; the two quads take alternate pairs, each with its own S / O accumulators in TMEM
//                     and its own P tile: tcgen05.ld of the row's 64 scores -> + rel-pos bias -> + shift mask -> max / ex2 /
//                     sum in registers -> normalised P in bf16 to the swizzled A-operand tile -> tcgen05.ld of O -> bf16 ->
//                     scatter to token order.
// The block-diagonal P tile ([128 x 128], zero off the two 64 x 64 diagonal blocks) costs 24 KB instead of 32: its two
// K blocks overlap in a shared zero region.  Same math as window_attn_fwd7_kernel (log2-domain scores, P normalised in
// fp32 then rounded to bf16, natural-log LSE saved for the backward).
#pragma once
#include <cuda.h>
#include "wa_common.cuh"

namespace wa {
namespace tc {

constexpr int ROWS = 128;                 // 2 windows x 64 slots
constexpr int TILE_B = ROWS * 128;        // bytes of one operand tile: 128 rows x 128 B
constexpr int STAGE_B = 2 * TILE_B;       // [Q | -] and [K | V]
constexpr int NSTAGE = 4;
constexpr int GDEPTH = 2;               // a pair's copies are awaited GDEPTH pairs after they were issued
constexpr int P_B = 3 * 8192;             // block-diagonal P: [data0 | zero | data1], K block kb starts at kb * 8 KB
constexpr int BIAS_LD = 68;               // floats per row of the staged rel-pos bias (conflict-free float4 reads)
constexpr int NTHREADS = 32 * 13;         // warps 0-7 rows (2 quads), 8-11 gather, 12 MMA
constexpr int TMEM_COLS = 512;
constexpr int BUF_COLS = 192;             // per quad: S 128 columns + O 64 columns

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
// Bounded wait (~2 s: a mis-programmed pipeline traps instead of hanging the GPU).  The try_wait carries a suspend-time hint
// and failed tries back off with nanosleep: up to ten of the CTA's thirteen warps are waiting at any time, and a hot spin
// loop in each of them took more issue slots than the softmax itself (ncu: 400 try_wait iterations per window pair).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(a), "r"(parity), "r"(1000u)
        : "memory");
    if (done) return;
    __nanosleep(it < 4 ? 20 : 100);
    if ((it & 255u) == 255u) {
      const uint64_t t = global_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 2000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}
// SmemDescriptor, 128B swizzle: K-major (rows of 128 B, 8-row atoms 1024 B apart) or MN-major ([K rows][64 MN] atoms)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, bool mn_major) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(mn_major ? (8192 >> 4) : 1) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }


// ---- output through bulk tensor stores ---------------------------------------------------------------------------------
// One (window, tensor) box of an output: 49 slots x 32 channels, staged in shared memory (64-byte rows, 64B swizzle) and
// written by ONE bulk tensor store: rows of padded slots fall outside the [B, H, W, ch] tensor and are clipped.  A window
// that wraps around the image (last window row / column of a shifted block) is staged as 2 or 4 dense sub-boxes, one per
// wrapped copy, each with its own store and box shape: [7 - shift | shift] columns x [7 - shift | shift] rows (a store whose
// box STARTS at a negative coordinate raises an illegal-instruction error on sm_100, measured; only the upper bound is
// clipped).  Per-thread 16-byte stores of token rows, the alternative, cost four times the rest of the pair's work: the LSU
// handles one 128-byte line per request and such an instruction touches 32 lines.
struct OutMaps { CUtensorMap m[9]; };   // box (32 ch, bx, by, 1): m[iy * 3 + ix], bx / by in {7, 7 - shift, shift}
constexpr int OBOX_B = 3584;            // 49 * 64 = 3136, rounded up to the 512-byte period of the swizzle
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(map),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// row of slot (siy, six) inside its window's staged box; xw / yw: the window wraps around the image in x / in y
__device__ __forceinline__ int obox_row(int siy, int six, bool xw, bool yw, int shift) {
  const int n1 = 7 - shift;
  const int sx = xw && six >= n1, sy = yw && siy >= n1;
  const int bw = xw ? (sx ? shift : n1) : 7, bh = yw ? (sy ? shift : n1) : 7;
  return sy * n1 * 7 + sx * n1 * bh + (siy - (sy ? n1 : 0)) * bw + (six - (sx ? n1 : 0));
}
// the stores of window `win`'s staged box at shared address src (channel offset ch0): 1, 2 or 4 of them, one bulk group
template <bool SHIFT>
__device__ __forceinline__ void obox_store(const OutMaps& om, uint32_t src, int ch0, int win, const Geo& g) {
  const int n1 = 7 - g.shift;
  const int wx = win % g.nWx, t2 = win / g.nWx;
  const int wy = t2 % g.nWy, bb = t2 / g.nWy;
  const int x0 = wx * 7 + g.shift, y0 = wy * 7 + g.shift;   // x0 < Wp, y0 < Hp
  const int xw = SHIFT && wx == g.nWx - 1, yw = SHIFT && wy == g.nWy - 1;
  for (int sy = 0; sy <= yw; sy++)
    for (int sx = 0; sx <= xw; sx++) {
      const int bh = yw ? (sy ? g.shift : n1) : 7;
      const int srow = sy * n1 * 7 + sx * n1 * bh;
      tma_store_4d(&om.m[(yw ? 1 + sy : 0) * 3 + (xw ? 1 + sx : 0)], src + srow * 64, ch0, sx ? 0 : x0, sy ? 0 : y0, bb);
    }
}

struct Smem {
  // byte offsets from the 1024-aligned base
  static constexpr int STAGES = 0;                               // [NSTAGE][Q | KV]
  static constexpr int P = NSTAGE * STAGE_B;                     // [2 quads][24 KB]
  static constexpr int BIAS = P + 2 * P_B;                       // [64][68] fp32
  static constexpr int META = BIAS + 64 * BIAS_LD * 4;           // tok [NSTAGE][128] int, rid [NSTAGE][128] int
  static constexpr int BARS = META + 2 * NSTAGE * ROWS * 4;      // mbarriers + tmem ptr
  static constexpr int TOTAL = BARS + 256;
};
static size_t fwd7_tc_smem() { return (size_t)Smem::TOTAL + 1024; }

template <bool SHIFT>
__global__ void __launch_bounds__(NTHREADS, 1) window_attn_fwd7_tc_kernel(
    const bf16* __restrict__ qkv, const bf16* __restrict__ qkv_bias, const float* __restrict__ bexp,
    bf16* __restrict__ out, float* __restrict__ lse, Geo g, float scale, int nwin_total,
    const __grid_constant__ OutMaps om) {
  constexpr int WS = 7, NT = 49;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = base + Smem::STAGES;
  uint8_t* pbuf = base + Smem::P;
  float* bias_s = reinterpret_cast<float*>(base + Smem::BIAS);
  int* tokb = reinterpret_cast<int*>(base + Smem::META);  // [3][128]
  int* ridb = tokb + NSTAGE * ROWS;                       // [3][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + Smem::BARS);
  uint64_t* full_qkv = bars;                 // [NSTAGE] count 128 (gather threads)
  uint64_t* empty_qkv = bars + NSTAGE;       // [NSTAGE] count 1   (MMA commit after P V)
  uint64_t* s_full = bars + 2 * NSTAGE;      // [2] count 1   (MMA commit)
  uint64_t* s_free = s_full + 2;             // [2] count 4   (row warps of the quad)
  uint64_t* p_full = s_full + 4;             // [2] count 4
  uint64_t* o_full = s_full + 6;             // [2] count 1   (MMA commit)
  uint64_t* o_free = s_full + 8;             // [2] count 4
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(s_full + 10);

  const int h = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npairs = (nwin_total + 1) >> 1;
  const int n_items = (int)blockIdx.y < npairs ? (npairs - 1 - (int)blockIdx.y) / (int)gridDim.y + 1 : 0;

  // one-time setup: zero every operand tile (unused half rows and the shared zero region of P feed the MMAs), stage this
  // head's rel-pos bias, barriers, TMEM
  for (int i = threadIdx.x; i < (NSTAGE * STAGE_B + 2 * P_B) / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(base)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < 64 * 16; i += NTHREADS) {
    const int row = i >> 4, c4 = (i & 15) * 4;
    *reinterpret_cast<float4*>(bias_s + row * BIAS_LD + c4) =
        __ldg(reinterpret_cast<const float4*>(bexp + (long long)h * 4096 + (row < NT ? row : 0) * 64 + c4));
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; i++) { mbar_init(&full_qkv[i], 128); mbar_init(&empty_qkv[i], 1); }
    for (int i = 0; i < 2; i++) {
      mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&o_free[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // the zero fill is visible to the tensor-core (async) proxy
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= 8 && warp < 12) {
    // ===================== gather warps =====================
    const int t = threadIdx.x - 256;
    const int c16 = t & 3;  // 4 adjacent lanes cover one 64-byte (slot, q|k|v) segment
    uint4 bchunk[3];
#pragma unroll
    for (int part = 0; part < 3; part++)
      bchunk[part] = __ldg(reinterpret_cast<const uint4*>(qkv_bias + part * g.C + h * HD + c16 * 8));
    // per-thread constants: the 4 rows this thread serves (same slot geometry for every pair)
    int rr[4], iy[4], ix[4];
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      rr[kk] = (t >> 2) + 32 * kk;
      const int i = rr[kk] & 63;
      iy[kk] = i / WS;
      ix[kk] = i - iy[kk] * WS;
    }
    // token row of row kk of window pair `pair` (-1 padded slot, -2 no such slot) and its shift region; (bb, wy, wx) of the
    // pair's two windows are computed once per pair, not per row
    // (image, window row, window column) of a window index, advanced INCREMENTALLY from pair to pair: the persistent loop
    // strides by a constant number of windows, so the div / mod by run-time values happens once per kernel, not per pair
    struct WinPos { int bb, wy, wx; };
    struct PairGeo { int bb[2], wy[2], wx[2]; bool ok[2]; };
    auto from_index = [&](int win) {
      WinPos q;
      q.wx = win % g.nWx;
      const int t2 = win / g.nWx;
      q.wy = t2 % g.nWy;
      q.bb = t2 / g.nWy;
      return q;
    };
    auto advance = [&](WinPos& q, const WinPos& sft) {
      q.wx += sft.wx;
      int cy = q.wx >= g.nWx ? 1 : 0;
      q.wx -= cy ? g.nWx : 0;
      q.wy += sft.wy + cy;
      cy = q.wy >= g.nWy ? 1 : 0;
      q.wy -= cy ? g.nWy : 0;
      q.bb += sft.bb + cy;
    };
    const WinPos wstep = from_index(2 * (int)gridDim.y), wone = {0, 0, 1};
    WinPos wcur = from_index(2 * (int)blockIdx.y), wpf = from_index(2 * ((int)blockIdx.y + NSTAGE * (int)gridDim.y));
    auto pair_geo = [&](const WinPos& w0, int pair) {
      PairGeo pg;
      WinPos w1 = w0;
      advance(w1, wone);
      pg.bb[0] = w0.bb; pg.wy[0] = w0.wy; pg.wx[0] = w0.wx; pg.ok[0] = 2 * pair < nwin_total;
      pg.bb[1] = w1.bb; pg.wy[1] = w1.wy; pg.wx[1] = w1.wx; pg.ok[1] = 2 * pair + 1 < nwin_total;
      return pg;
    };
    auto slot = [&](const PairGeo& pg, int kk, int& tk, int& rd) {
      const int w = rr[kk] >> 6;
      tk = -2; rd = 0;
      if (pg.ok[w] && iy[kk] < WS) {
        const int ry = pg.wy[w] * WS + iy[kk], rx = pg.wx[w] * WS + ix[kk];
        int py = ry + g.shift, px = rx + g.shift;
        if (py >= g.Hp) py -= g.Hp;
        if (px >= g.Wp) px -= g.Wp;
        tk = (py < g.H && px < g.W) ? (pg.bb[w] * g.H + py) * g.W + px : -1;
        if (g.shift > 0) {
          const int ay = (ry >= g.Hp - WS) + (ry >= g.Hp - g.shift);
          const int ax = (rx >= g.Wp - WS) + (rx >= g.Wp - g.shift);
          rd = ay * 3 + ax;
          // windows that wrap around the image in x / in y (flags for the row threads' output staging; the same for every
          // slot of a window, so the region comparisons are unaffected)
          if (pg.wx[w] == g.nWx - 1) rd |= 16;
          if (pg.wy[w] == g.nWy - 1) rd |= 32;
        }
      }
    };
    const int gdepth = (g.dbg & 128) ? 1 : GDEPTH;   // (A/B knob: publish a pair one gather iteration after its copies)
    for (int it = 0; it < n_items + gdepth; it++) {
      // FIRST publish the pair issued GDEPTH iterations ago (its copies have landed), THEN wait for a free stage: the stage
      // this iteration needs is released by MMAs that themselves wait for that publication (circular otherwise)
      if (it >= gdepth) {
        if (gdepth == 1) cp_async_wait<0>(); else cp_async_wait<GDEPTH - 1>();
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> visible to tcgen05.mma
        mbar_arrive(&full_qkv[(it - gdepth) % NSTAGE]);
      }
      if (it < n_items) {
        const int st_i = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        const int pair = blockIdx.y + it * gridDim.y;
        if (it + NSTAGE < n_items && c16 < 3 && (g.dbg & 64)) {  // L2 prefetch of the pair NSTAGE ahead (ESVIT_ATTN_DBG=64;
          // off: a prefetch instruction with 32 distinct lines costs the gather warp ~400 cycles, measured in the backward)
          const PairGeo pf = pair_geo(wpf, pair + NSTAGE * (int)gridDim.y);
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            int tk, rd;
            slot(pf, kk, tk, rd);
            if (tk >= 0) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(qkv + (long long)tk * 3 * g.C + c16 * g.C + h * HD));
          }
        }
        const PairGeo pg = pair_geo(wcur, pair);
        advance(wcur, wstep);
        advance(wpf, wstep);
        mbar_wait(&empty_qkv[st_i], ph ^ 1);
        uint8_t* stq = stages + st_i * STAGE_B;
        uint8_t* stkv = stq + TILE_B;
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          const int r = rr[kk];
          int tk, rd;
          slot(pg, kk, tk, rd);
          const int sw = r & 7;
          uint8_t* dq = stq + r * 128 + ((c16 ^ sw) * 16);
          uint8_t* dk = stkv + r * 128 + ((c16 ^ sw) * 16);
          uint8_t* dv = stkv + r * 128 + (((4 + c16) ^ sw) * 16);
          if (tk == -1) {                            // padded slot: the qkv bias (the reference pads norm1's output with zeros)
            *reinterpret_cast<uint4*>(dq) = bchunk[0];
            *reinterpret_cast<uint4*>(dk) = bchunk[1];
            *reinterpret_cast<uint4*>(dv) = bchunk[2];
          } else {
            const bf16* src = qkv + (long long)(tk >= 0 ? tk : 0) * 3 * g.C + h * HD + c16 * 8;
            const int nbytes = tk >= 0 ? 16 : 0;     // slots >= 49 / missing second window: zero fill
            cp_async16(dq, src, nbytes);
            cp_async16(dk, src + g.C, nbytes);
            cp_async16(dv, src + 2 * g.C, nbytes);
          }
          if (c16 == 0) {
            tokb[st_i * ROWS + r] = tk;
            ridb[st_i * ROWS + r] = rd;
          }
        }
      }
      cp_async_commit();   // (an empty group past the last pair keeps the group arithmetic uniform)
    }
  } else if (warp == 12) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc(128, 128, false);
      const uint32_t idesc_o = make_idesc(128, 64, true);
      // descriptors of stage 0 / quad 0; other stages and quads differ only in the 14-bit address field (units of 16 B)
      const uint64_t dq0 = make_desc(smem_u32(stages), false), dk0 = make_desc(smem_u32(stages) + TILE_B, false);
      const uint64_t dv0 = make_desc(smem_u32(stages) + TILE_B, true), dp0 = make_desc(smem_u32(pbuf), false);
      for (int it = 0; it <= n_items; it++) {
        if (it < n_items) {  // S(it) = Q K^T
          const int b = it & 1, st_i = it % NSTAGE;
          mbar_wait(&full_qkv[st_i], (it / NSTAGE) & 1);
          mbar_wait(&s_free[b], ((it >> 1) & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint64_t so = (uint64_t)(st_i * (STAGE_B >> 4));
#pragma unroll
          for (int k = 0; k < 2; k++) umma(tmem_base + b * BUF_COLS, dq0 + so + 2 * k, dk0 + so + 2 * k, idesc_s, k);
          umma_commit(&s_full[b]);
        }
        if (it > 0) {        // O(it-1) = P [K|V]
          const int j = it - 1, b = j & 1, st_j = j % NSTAGE;
          const uint32_t ph = (j >> 1) & 1;
          mbar_wait(&p_full[b], ph);
          mbar_wait(&o_free[b], ph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint64_t pa = dp0 + (uint64_t)(b * (P_B >> 4)), va = dv0 + (uint64_t)(st_j * (STAGE_B >> 4));
#pragma unroll
          for (int k = 0; k < 8; k++)   // A: K block k>>2 at +8 KB, 16 keys = 32 B along K; B: 16 key rows of [K|V] = 2 KB
            umma(tmem_base + b * BUF_COLS + 128, pa + (uint64_t)((k >> 2) * (8192 >> 4) + 2 * (k & 3)), va + (uint64_t)(128 * k), idesc_o, k);
          umma_commit(&o_full[b]);
          umma_commit(&empty_qkv[st_j]);  // the stage's tiles are free once these MMAs have read them
        }
      }
    }
  } else {
    // ===================== row warps: two quads, thread = query row =====================
    const int quad = warp >> 2, qw = warp & 3;
    const int r = qw * 32 + lane;              // 0..127 = TMEM lane
    const int w = r >> 6, i = r & 63;
    const uint32_t taddr = tmem_base + ((uint32_t)(qw * 32) << 16) + quad * BUF_COLS;
    const float c = scale * LOG2E;
    const float* brow = bias_s + i * BIAS_LD;
    uint8_t* prow = pbuf + quad * P_B + w * 8192 + r * 128;
    uint8_t* obox = pbuf + quad * P_B + w * OBOX_B;   // this window's output box: in data half 0 of the quad's P tile
    const int siy = i / WS, six = i - siy * WS;
    const bool issuer = lane == 0 && qw < 2;          // issues the bulk tensor stores of window qw of each pair
    for (int it = quad; it < n_items; it += 2) {
      const int st_i = it % NSTAGE;
      const uint32_t ph = (it >> 1) & 1;
      const int pair = blockIdx.y + it * gridDim.y;
      mbar_wait(&full_qkv[st_i], (it / NSTAGE) & 1);   // tok / rid of this stage are written
      const int tok = tokb[st_i * ROWS + r];
      int rid_r = 0;
      if (SHIFT) rid_r = ridb[st_i * ROWS + r];
      mbar_wait(&s_full[quad], ph);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      uint32_t v[64];
      tmem_ld32(taddr + w * 64, v);
      tmem_ld32(taddr + w * 64 + 32, v + 32);
      tmem_ld_wait();
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[quad]);
      float x[64];
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int j4 = 0; j4 < 16; j4++) {
        const float4 b4 = *reinterpret_cast<const float4*>(brow + j4 * 4);
        x[4 * j4] = fmaf(__uint_as_float(v[4 * j4]), c, b4.x);
        x[4 * j4 + 1] = fmaf(__uint_as_float(v[4 * j4 + 1]), c, b4.y);
        x[4 * j4 + 2] = fmaf(__uint_as_float(v[4 * j4 + 2]), c, b4.z);
        x[4 * j4 + 3] = fmaf(__uint_as_float(v[4 * j4 + 3]), c, b4.w);
        if (SHIFT) {
          if (4 * j4 < NT) {
            const int4 rc = *reinterpret_cast<const int4*>(ridb + st_i * ROWS + w * 64 + 4 * j4);
            if (rc.x != rid_r) x[4 * j4] += -100.f * LOG2E;
            if (rc.y != rid_r) x[4 * j4 + 1] += -100.f * LOG2E;
            if (rc.z != rid_r) x[4 * j4 + 2] += -100.f * LOG2E;
            if (rc.w != rid_r) x[4 * j4 + 3] += -100.f * LOG2E;
          }
        }
        m0 = fmaxf(m0, x[4 * j4]); m1 = fmaxf(m1, x[4 * j4 + 1]); m2 = fmaxf(m2, x[4 * j4 + 2]); m3 = fmaxf(m3, x[4 * j4 + 3]);
      }
      const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int j4 = 0; j4 < 16; j4++) {
        x[4 * j4] = ex2(x[4 * j4] - m); x[4 * j4 + 1] = ex2(x[4 * j4 + 1] - m);
        x[4 * j4 + 2] = ex2(x[4 * j4 + 2] - m); x[4 * j4 + 3] = ex2(x[4 * j4 + 3] - m);
        s0 += x[4 * j4]; s1 += x[4 * j4 + 1]; s2 += x[4 * j4 + 2]; s3 += x[4 * j4 + 3];
      }
      const float s = (s0 + s1) + (s2 + s3);
      const float inv = __fdividef(1.f, s);
      // normalised P (fp32 -> bf16, as the reference under autocast) into this window's diagonal block of the A-operand tile
      // (the quad's previous P V has completed: its o_full was awaited in the previous iteration's epilogue; the bulk stores
      // of its output, staged in the same tile, have finished reading: awaited by the issuing threads before the barrier)
      if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
      asm volatile("bar.sync %0, 128;\n" ::"r"(quad + 1) : "memory");
#pragma unroll
      for (int ch = 0; ch < 8; ch++) {
        float p8[8];
#pragma unroll
        for (int e = 0; e < 8; e++) p8[e] = x[ch * 8 + e] * inv;
        *reinterpret_cast<bf16x8*>(prow + ((ch ^ (r & 7)) * 16)) = pack8(p8);
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[quad]);
      const int win = 2 * pair + w;
      if (win < nwin_total && i < NT) lse[((long long)win * g.nH + h) * NT + i] = (m + lg2(s)) * LN2;
      // epilogue: O -> bf16 -> token order (the other quad's softmax runs meanwhile)
      mbar_wait(&o_full[quad], ph);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      uint32_t o[32];
      tmem_ld32(taddr + 128 + 32, o);            // columns 32..63 of P [K|V] = P V
      tmem_ld_wait();
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[quad]);
      if (i < NT) {
        const int orow = SHIFT ? obox_row(siy, six, (rid_r & 16) != 0, (rid_r & 32) != 0, g.shift) : i;
        uint8_t* dst = obox + orow * 64;
        const int sw = (orow >> 1) & 3;
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
          float f8[8];
#pragma unroll
          for (int e = 0; e < 8; e++) f8[e] = __uint_as_float(o[ch * 8 + e]);
          *reinterpret_cast<bf16x8*>(dst + ((ch ^ sw) * 16)) = pack8(f8);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      asm volatile("bar.sync %0, 128;\n" ::"r"(quad + 1) : "memory");
      if (issuer) {
        const int win2 = 2 * pair + qw;
        if (win2 < nwin_total) obox_store<SHIFT>(om, smem_u32(pbuf + quad * P_B + qw * OBOX_B), h * HD, win2, g);
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS));
}

}  // namespace tc
}  // namespace wa
