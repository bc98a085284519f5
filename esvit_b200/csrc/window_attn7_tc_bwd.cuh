// ws = 7 (shifted-)window attention core, BACKWARD, on tcgen05 tensor cores with TMEM accumulators.
//
// Same skeleton as window_attn7_tc.cuh: one persistent CTA per SM serves ONE head and walks over PAIRS of windows (2 x 64
// slots = the 128 rows of one UMMA tile, thread = row).  Per pair, FIVE GEMMs run on the tensor cores:
//
//   S  = Q K^T      dP = dO V^T                         M=128 N=128 K=32   (tiles [Q | dO] and [K | V]: a slot's q / dO and
//                                                                            k / v share one 128-byte row; K-major)
//   row quad:  P = exp2(S c + bias + mask - lse),  dS = P (dP - D),  D = rowsum(dO o O);  P, dS -> bf16 block-diagonal tiles
//   dQ = dS K       A = dS K-major,   B = [K|V] MN-major (N = 32: the K half)        M=128 N=32 K=128
//   dK = dS^T Q     A = dS MN-major,  B = [Q|dO] MN-major (N = 32: the Q half)       M=128 N=32 K=128
//   dV = P^T dO     A = P  MN-major,  B = [Q|dO] MN-major (N = 64: columns 32..63)   M=128 N=64 K=128
//
// i.e. the transposes the mma.sync kernel obtains by recomputing S^T / dP^T are free here: the SAME shared-memory tile is
// read K-major or MN-major by the UMMA descriptor.  dQ / dK / dV re-use the TMEM columns of S / dP (consumed by then), so
// a quad needs 256 columns and two quads alternate pairs.  The block-diagonal P / dS tiles cost 24 KB each (their two K
// blocks / M atoms overlap in a shared zero region).
// Gradients of the small parameters stay in this kernel: the rel-pos-bias gradient is accumulated in registers per row
// thread over all windows (flushed through shared-memory bins once per CTA), the qkv-bias gradient = column sums of
// dQ / dK / dV by a 31-shuffle butterfly per warp and tensor, accumulated in one register per lane.
#pragma once
#include "wa_common.cuh"
#include "window_attn7_tc.cuh"

namespace wa {
namespace tcb {

using tc::elect_one;
using tc::make_desc;
using tc::mbar_arrive;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;
using tc::tmem_ld32;
using tc::tmem_ld_wait;
using tc::umma;
using tc::umma_commit;

constexpr int ROWS = 128;
constexpr int TILE_B = ROWS * 128;        // 16 KB
constexpr int STAGE_B = 2 * TILE_B;       // [Q | dO] and [K | V]
constexpr int NSTAGE = 3;
constexpr int GDEPTH = 2;               // a pair's copies are awaited GDEPTH pairs after they were issued
constexpr int PD_B = 3 * 8192;            // block-diagonal P or dS: [data0 | zero | data1]
constexpr int BIAS_LD = 68;
constexpr int NTHREADS = 32 * 11;         // warps 0-7 rows (2 quads), 8-9 gather, 10 MMA
constexpr int TMEM_COLS = 512;
constexpr int BUF_COLS = 256;             // per quad: S 128 + dP 128, re-used as dQ 32 | dK 32 | dV 64

// InstrDescriptor with both operand majors
__device__ __forceinline__ uint32_t make_idesc2(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

struct Smem {
  static constexpr int STAGES = 0;                               // [3][ [Q|dO] | [K|V] ]
  static constexpr int PD = NSTAGE * STAGE_B;                    // [2 quads][P 24 KB | dS 24 KB]
  static constexpr int BIAS = PD + 2 * 2 * PD_B;                 // [64][68] fp32
  static constexpr int BINS = BIAS + 64 * BIAS_LD * 4;           // [169 + 3] fp32 rel-pos-bias gradient bins
  static constexpr int META = BINS + 176 * 4;                    // tok [3][128] int, rid [3][128] int
  static constexpr int BARS = META + 2 * NSTAGE * ROWS * 4;
  static constexpr int TOTAL = BARS + 256;
};
static size_t bwd7_tc_smem() { return (size_t)Smem::TOTAL + 1024; }

// column sums of a [32 rows (lanes)] x [32 columns (v[0..31])] tile: after the butterfly lane l holds the sum of column
// bitrev5(l)... the mapping is irrelevant as long as the flush uses the same one: lane l ends with column col_of_lane(l).
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
  // step s: partners differ in bit s of the lane id; the lane with bit = 0 keeps the lower half of the live columns
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; j++) {
      const float mine = upper ? v[j + half] : v[j];
      const float send = upper ? v[j] : v[j + half];
      v[j] = mine + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}
// the column whose sum lane l holds after warp_colsum32: bit `half` of the lane selects the upper half at that step
__device__ __forceinline__ int colsum_col_of_lane(int lane) { return lane & 31; }

template <bool SHIFT>
__global__ void __launch_bounds__(NTHREADS, 1) window_attn_bwd7_tc_kernel(
    const bf16* __restrict__ qkv, const bf16* __restrict__ qkv_bias, const float* __restrict__ bexp,
    const bf16* __restrict__ out, const bf16* __restrict__ dout, const float* __restrict__ lse,
    bf16* __restrict__ dqkv, float* __restrict__ dbias_table, float* __restrict__ dqkv_bias, Geo g, float scale,
    int nwin_total) {
  constexpr int WS = 7, NT = 49, NB = 169;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = base + Smem::STAGES;
  uint8_t* pdbuf = base + Smem::PD;
  float* bias_s = reinterpret_cast<float*>(base + Smem::BIAS);
  float* bins = reinterpret_cast<float*>(base + Smem::BINS);
  int* tokb = reinterpret_cast<int*>(base + Smem::META);
  int* ridb = tokb + NSTAGE * ROWS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + Smem::BARS);
  uint64_t* full_in = bars;         // [3] count 64  (gather threads)
  uint64_t* empty_in = bars + 3;    // [3] count 1   (MMA commit after the second GEMM group)
  uint64_t* s_full = bars + 6;      // [2] count 1   (S and dP complete)
  uint64_t* s_free = bars + 8;      // [2] count 4   (row warps have loaded S / dP)
  uint64_t* pd_full = bars + 10;    // [2] count 4   (P and dS tiles written)
  uint64_t* g_full = bars + 12;     // [2] count 1   (dQ / dK / dV complete)
  uint64_t* g_free = bars + 14;     // [2] count 4   (row warps have loaded dQ / dK / dV: the accumulator columns are free)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);

  const int h = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npairs = (nwin_total + 1) >> 1;
  const int n_items = (int)blockIdx.y < npairs ? (npairs - 1 - (int)blockIdx.y) / (int)gridDim.y + 1 : 0;

  for (int i = threadIdx.x; i < (NSTAGE * STAGE_B + 4 * PD_B) / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(base)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < 64 * 16; i += NTHREADS) {
    const int row = i >> 4, c4 = (i & 15) * 4;
    *reinterpret_cast<float4*>(bias_s + row * BIAS_LD + c4) =
        __ldg(reinterpret_cast<const float4*>(bexp + (long long)h * 4096 + (row < NT ? row : 0) * 64 + c4));
  }
  for (int i = threadIdx.x; i < 176; i += NTHREADS) bins[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; i++) { mbar_init(&full_in[i], 64); mbar_init(&empty_in[i], 1); }
    for (int i = 0; i < 2; i++) {
      mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4); mbar_init(&pd_full[i], 4); mbar_init(&g_full[i], 1); mbar_init(&g_free[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 10) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 8 || warp == 9) {
    // ===================== gather warps: 64 threads, 4 lanes per 64-byte segment, 8 rows per thread =====================
    const int t = threadIdx.x - 256;
    const int c16 = t & 3;
    uint4 bchunk[3];
#pragma unroll
    for (int part = 0; part < 3; part++)
      bchunk[part] = __ldg(reinterpret_cast<const uint4*>(qkv_bias + part * g.C + h * HD + c16 * 8));
    int iy[8], ix[8];   // slot geometry of this thread's 8 rows r = (t >> 2) + 16 kk (the same for every pair)
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
      const int i = ((t >> 2) + 16 * kk) & 63;
      iy[kk] = i / WS;
      ix[kk] = i - iy[kk] * WS;
    }
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
    auto slot = [&](const PairGeo& pg, int kk, int& tk, int& rd) {  // -1 padded slot, -2 no such slot
      const int w = kk >> 2;  // rows 16 kk + (t >> 2): kk < 4 -> window 0, else window 1
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
        }
      }
    };
    for (int it = 0; it < n_items + GDEPTH; it++) {
      // FIRST publish the pair issued GDEPTH iterations ago (its copies have landed), THEN wait for a free stage: the stage
      // this iteration needs is released by GEMMs that themselves wait for that publication (circular otherwise)
      if (it >= GDEPTH) {
        cp_async_wait<GDEPTH - 1>();
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        mbar_arrive(&full_in[(it - GDEPTH) % NSTAGE]);
      }
      if (it < n_items) {
        const int st_i = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        const int pair = blockIdx.y + it * gridDim.y;
        if (it + NSTAGE < n_items) {  // L2 prefetch of the pair three ahead (q / k / v segments, dO and O rows)
          const PairGeo pf = pair_geo(wpf, pair + NSTAGE * (int)gridDim.y);
#pragma unroll
          for (int kk = 0; kk < 8; kk++) {
            int tk, rd;
            slot(pf, kk, tk, rd);
            if (tk >= 0) {
              if (c16 < 3) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(qkv + (long long)tk * 3 * g.C + c16 * g.C + h * HD));
              else {
                asm volatile("prefetch.global.L2 [%0];\n" ::"l"(dout + (long long)tk * g.C + h * HD));
                asm volatile("prefetch.global.L2 [%0];\n" ::"l"(out + (long long)tk * g.C + h * HD));
              }
            }
          }
        }
        const PairGeo pg = pair_geo(wcur, pair);
        advance(wcur, wstep);
        advance(wpf, wstep);
        mbar_wait(&empty_in[st_i], ph ^ 1);
        uint8_t* t1 = stages + st_i * STAGE_B;   // [Q | dO]
        uint8_t* t2 = t1 + TILE_B;               // [K | V]
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
          const int r = (t >> 2) + 16 * kk;
          int tk, rd;
          slot(pg, kk, tk, rd);
          const int sw = r & 7;
          uint8_t* dq = t1 + r * 128 + ((c16 ^ sw) * 16);
          uint8_t* dd = t1 + r * 128 + (((4 + c16) ^ sw) * 16);
          uint8_t* dk = t2 + r * 128 + ((c16 ^ sw) * 16);
          uint8_t* dv = t2 + r * 128 + (((4 + c16) ^ sw) * 16);
          if (tk == -1) {          // padded slot: q / k / v = the qkv bias, its output row is cropped away: dO = 0
            *reinterpret_cast<uint4*>(dq) = bchunk[0];
            *reinterpret_cast<uint4*>(dk) = bchunk[1];
            *reinterpret_cast<uint4*>(dv) = bchunk[2];
            *reinterpret_cast<uint4*>(dd) = make_uint4(0u, 0u, 0u, 0u);
          } else {
            const long long trow = tk >= 0 ? tk : 0;
            const bf16* src = qkv + trow * 3 * g.C + h * HD + c16 * 8;
            const int nbytes = tk >= 0 ? 16 : 0;
            cp_async16(dq, src, nbytes);
            cp_async16(dk, src + g.C, nbytes);
            cp_async16(dv, src + 2 * g.C, nbytes);
            cp_async16(dd, dout + trow * g.C + h * HD + c16 * 8, nbytes);
          }
          if (c16 == 0) {
            tokb[st_i * ROWS + r] = tk;
            ridb[st_i * ROWS + r] = rd;
          }
        }
      }
      cp_async_commit();
    }
  } else if (warp == 10) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t id_s = make_idesc2(128, 128, false, false);
      const uint32_t id_dq = make_idesc2(128, 32, false, true);
      const uint32_t id_dk = make_idesc2(128, 32, true, true);
      const uint32_t id_dv = make_idesc2(128, 64, true, true);
      // descriptors of stage 0 / quad 0; other stages and quads differ only in the 14-bit address field (units of 16 B)
      const uint32_t s0 = smem_u32(stages), p0 = smem_u32(pdbuf);
      const uint64_t t1k = make_desc(s0, false), t2k = make_desc(s0 + TILE_B, false);      // [Q|dO], [K|V] K-major
      const uint64_t t1m = make_desc(s0, true), t2m = make_desc(s0 + TILE_B, true);        // ... MN-major
      const uint64_t pm = make_desc(p0, true), dsk = make_desc(p0 + PD_B, false), dsm = make_desc(p0 + PD_B, true);
      for (int it = 0; it <= n_items; it++) {
        if (it < n_items) {  // S(it), dP(it)
          const int b = it & 1, st_i = it % NSTAGE;
          mbar_wait(&full_in[st_i], (it / NSTAGE) & 1);
          mbar_wait(&g_free[b], ((it >> 1) & 1) ^ 1);   // the quad has drained dQ / dK / dV of its previous pair
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint64_t so = (uint64_t)(st_i * (STAGE_B >> 4));
#pragma unroll
          for (int k = 0; k < 2; k++) umma(tmem_base + b * BUF_COLS, t1k + so + 2 * k, t2k + so + 2 * k, id_s, k);                 // Q K^T
#pragma unroll
          for (int k = 0; k < 2; k++) umma(tmem_base + b * BUF_COLS + 128, t1k + so + 4 + 2 * k, t2k + so + 4 + 2 * k, id_s, k);   // dO V^T
          umma_commit(&s_full[b]);
        }
        if (it > 0) {        // dQ, dK, dV of pair it-1
          const int j = it - 1, b = j & 1, st_j = j % NSTAGE;
          const uint32_t ph = (j >> 1) & 1;
          mbar_wait(&pd_full[b], ph);                   // implies the quad has loaded S / dP: their columns are free
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint64_t so = (uint64_t)(st_j * (STAGE_B >> 4)), po = (uint64_t)(b * ((2 * PD_B) >> 4));
          const uint32_t col = tmem_base + b * BUF_COLS;
#pragma unroll
          for (int k = 0; k < 8; k++)   // dQ = dS K : A K-major (K block k>>2 at +8 KB), B = [K|V] MN-major rows 16k..
            umma(col, dsk + po + (uint64_t)((k >> 2) * (8192 >> 4) + 2 * (k & 3)), t2m + so + (uint64_t)(128 * k), id_dq, k);
#pragma unroll
          for (int k = 0; k < 8; k++)   // dK = dS^T Q : A = dS MN-major (query rows 16k..), B = [Q|dO] MN-major
            umma(col + 32, dsm + po + (uint64_t)(128 * k), t1m + so + (uint64_t)(128 * k), id_dk, k);
#pragma unroll
          for (int k = 0; k < 8; k++)   // [. | dV] = P^T [Q|dO]
            umma(col + 64, pm + po + (uint64_t)(128 * k), t1m + so + (uint64_t)(128 * k), id_dv, k);
          umma_commit(&g_full[b]);
          umma_commit(&empty_in[st_j]);
        }
      }
    }
  } else {
    // ===================== row warps: two quads, thread = row (query row in phase 1, q / k / v row in phase 2) =========
    const int quad = warp >> 2, qw = warp & 3;
    const int r = qw * 32 + lane;
    const int w = r >> 6, i = r & 63;
    const uint32_t taddr = tmem_base + ((uint32_t)(qw * 32) << 16) + quad * BUF_COLS;
    const float c = scale * LOG2E;
    const float* brow = bias_s + i * BIAS_LD;
    uint8_t* prow = pdbuf + quad * 2 * PD_B + w * 8192 + r * 128;   // P row; dS row at + PD_B
    float dsacc[64];
#pragma unroll
    for (int j = 0; j < 64; j++) dsacc[j] = 0.f;
    float csum[3] = {0.f, 0.f, 0.f};  // per-lane column-sum accumulators of dQ, dK, dV (column colsum_col_of_lane(lane))
    for (int it = quad; it < n_items; it += 2) {
      const int st_i = it % NSTAGE;
      const uint32_t ph = (it >> 1) & 1;
      const int pair = blockIdx.y + it * gridDim.y;
      mbar_wait(&full_in[st_i], (it / NSTAGE) & 1);
      const int tok = tokb[st_i * ROWS + r];
      int rid_r = 0;
      if (SHIFT) rid_r = ridb[st_i * ROWS + r];
      // D = rowsum(dO o O) and the row's LSE, straight from global memory (L2: prefetched by the gather warps)
      float Dr = 0.f, l2 = INFINITY;  // rows that do not exist: P = exp2(-inf) = 0
      if (tok >= 0) {
        const uint4* po = reinterpret_cast<const uint4*>(out + (long long)tok * g.C + h * HD);
        const uint4* pd = reinterpret_cast<const uint4*>(dout + (long long)tok * g.C + h * HD);
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
          float fo[8], fd[8];
          const uint4 uo = __ldg(po + ch), ud = __ldg(pd + ch);
          unpack8(*reinterpret_cast<const bf16x8*>(&uo), fo);
          unpack8(*reinterpret_cast<const bf16x8*>(&ud), fd);
#pragma unroll
          for (int e = 0; e < 8; e++) Dr = fmaf(fo[e], fd[e], Dr);
        }
      }
      // padded query rows (tok == -1) have dO = 0, hence dS = 0 and no dV contribution: P = 0 serves them too (and keeps
      // an undefined saved LSE of a skipped all-padding tile out of the arithmetic)
      if (tok >= 0) l2 = __ldg(lse + ((long long)(2 * pair + w) * g.nH + h) * NT + i) * LOG2E;
      mbar_wait(&s_full[quad], ph);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      // the quad's previous dQ / dK / dV GEMMs have completed (g_full awaited below in the previous iteration), so the
      // P / dS tiles may be overwritten
#pragma unroll
      for (int hh = 0; hh < 2; hh++) {
        uint32_t sv[32], dv[32];
        tmem_ld32(taddr + w * 64 + hh * 32, sv);
        tmem_ld32(taddr + 128 + w * 64 + hh * 32, dv);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
          float p8[8], d8[8];
          const float4 b0 = *reinterpret_cast<const float4*>(brow + hh * 32 + ch * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(brow + hh * 32 + ch * 8 + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          int rc[8];
          if (SHIFT) {
            const int4 r0 = *reinterpret_cast<const int4*>(ridb + st_i * ROWS + w * 64 + hh * 32 + ch * 8);
            const int4 r1 = *reinterpret_cast<const int4*>(ridb + st_i * ROWS + w * 64 + hh * 32 + ch * 8 + 4);
            rc[0] = r0.x; rc[1] = r0.y; rc[2] = r0.z; rc[3] = r0.w; rc[4] = r1.x; rc[5] = r1.y; rc[6] = r1.z; rc[7] = r1.w;
          }
#pragma unroll
          for (int e = 0; e < 8; e++) {
            const int j = hh * 32 + ch * 8 + e;
            float x = fmaf(__uint_as_float(sv[ch * 8 + e]), c, bb[e]) - l2;
            if (SHIFT) { if (j < NT && rc[e] != rid_r) x += -100.f * LOG2E; }
            const float pj = ex2(x);
            const float dsj = pj * (__uint_as_float(dv[ch * 8 + e]) - Dr);
            p8[e] = pj;
            d8[e] = dsj;
            dsacc[j] += dsj;
          }
          const int sw = ((hh * 4 + ch) ^ (r & 7)) * 16;
          *reinterpret_cast<bf16x8*>(prow + sw) = pack8(p8);
          *reinterpret_cast<bf16x8*>(prow + PD_B + sw) = pack8(d8);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(&s_free[quad]); mbar_arrive(&pd_full[quad]); }
      // phase 2: dQ / dK / dV of this pair -> bf16 -> token order; column sums -> qkv-bias gradient
      mbar_wait(&g_full[quad], ph);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      // one tensor at a time (32 live values): dQ at +0, dK at +32, dV at +96 (columns 32..63 of [. | dV])
#pragma unroll
      for (int part = 0; part < 3; part++) {
        uint32_t gr[32];
        tmem_ld32(taddr + (part == 0 ? 0 : (part == 1 ? 32 : 96)), gr);
        tmem_ld_wait();
        if (part == 2) {  // all three accumulators are in registers / consumed: the columns may be overwritten
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&g_free[quad]);
        }
        float f[32];
        const float sc = part == 2 ? 1.f : scale;
#pragma unroll
        for (int e = 0; e < 32; e++) f[e] = __uint_as_float(gr[e]) * sc;
        if (tok >= 0) {
          bf16* dst = dqkv + (long long)tok * 3 * g.C + part * g.C + h * HD;
#pragma unroll
          for (int ch = 0; ch < 4; ch++) *reinterpret_cast<bf16x8*>(dst + ch * 8) = pack8(f + ch * 8);
        }
        // qkv-bias gradient: column sums over ALL slots of the window (padded ones included; missing rows are 0)
        csum[part] += warp_colsum32(f, lane);
      }
    }
    // flush: rel-pos-bias gradient through the CTA's shared-memory bins, qkv-bias gradient straight to global memory
    if (i < NT) {
#pragma unroll
      for (int j = 0; j < 64; j++)
        if (j < NT) atomicAdd(&bins[bias_index<WS>(i, j)], dsacc[j]);
    }
    {
      const int col = colsum_col_of_lane(lane);
#pragma unroll
      for (int part = 0; part < 3; part++) atomicAdd(&dqkv_bias[part * g.C + h * HD + col], csum[part]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  for (int i = threadIdx.x; i < NB; i += NTHREADS) atomicAdd(&dbias_table[i * g.nH + h], bins[i]);
  if (warp == 10) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS));
}

}  // namespace tcb
}  // namespace wa
