// Shared device helpers of the window-attention kernels (mma.sync / ldmatrix wrappers, window-slot geometry).
#pragma once
#include "common.cuh"

namespace wa {

constexpr int HD = 32;  // head dim
constexpr int LD = 40;  // smem row stride (bf16 elements) of the q/k/v/dO tiles: 80 B rows -> conflict-free ldmatrix

struct Geo {
  int B, H, W, C, nH, shift, Hp, Wp, nWx, nWy;
  int dbg;  // profiling aid (ESVIT_ATTN_DBG): 1 = gather only the first window per CTA, 2 = skip the math
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// slot i of window (b, wy, wx) -> token row in the [B*H*W] map (or -1 for a padded slot) and shift-region id
template <int WS>
__device__ __forceinline__ void slot_info(const Geo& g, int b, int wy, int wx, int i, int& tok, int& rid) {
  const int iy = i / WS, ix = i - iy * WS;
  const int ry = wy * WS + iy, rx = wx * WS + ix;  // coordinates in the rolled, padded frame
  int py = ry + g.shift, px = rx + g.shift;          // rolled[r] = padded[(r + shift) mod Hp]
  if (py >= g.Hp) py -= g.Hp;
  if (px >= g.Wp) px -= g.Wp;
  tok = (py < g.H && px < g.W) ? (b * g.H + py) * g.W + px : -1;
  rid = 0;
  if (g.shift > 0) {
    const int ay = (ry >= g.Hp - WS) + (ry >= g.Hp - g.shift);
    const int ax = (rx >= g.Wp - WS) + (rx >= g.Wp - g.shift);
    rid = ay * 3 + ax;
  }
}

template <int WS>
struct Cfg {
  static constexpr int NT = WS * WS;            // tokens per window
  static constexpr int MT = (NT + 15) / 16;     // 16-row tiles
  static constexpr int KP = MT * 16;            // padded token count
  static constexpr int NT8 = KP / 8;            // 8-key tiles
  static constexpr int NW = (WS == 7) ? 4 : 7;  // warps
  static constexpr int NB = (2 * WS - 1) * (2 * WS - 1);
};

// column sums of a 16 x 32 fp32 accumulator tile (4 d-tiles x C-fragment) added to dst[32] in shared memory
__device__ __forceinline__ void colsum_to_smem(const float (&t)[4][4], float scale, float* dst, int lane) {
#pragma unroll
  for (int dt = 0; dt < 4; dt++) {
    float c0 = (t[dt][0] + t[dt][2]) * scale, c1 = (t[dt][1] + t[dt][3]) * scale;
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o);
      c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    }
    if (lane < 4) {
      atomicAdd(&dst[dt * 8 + lane * 2], c0);
      atomicAdd(&dst[dt * 8 + lane * 2 + 1], c1);
    }
  }
}

template <int WS>
__device__ __forceinline__ int bias_index(int i, int j) {
  const int yi = i / WS, xi = i - yi * WS, yj = j / WS, xj = j - yj * WS;
  return (yi - yj + WS - 1) * (2 * WS - 1) + (xi - xj + WS - 1);
}


constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 16-byte async copy global -> shared; src_bytes = 0 zero-fills the destination (nothing is read)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

}  // namespace wa
