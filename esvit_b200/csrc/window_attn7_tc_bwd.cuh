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
// dQ / dK / dV by a 16-shuffle butterfly per warp and tensor, accumulated in one register per lane.
// TWO threads serve each row (16 row warps): the exp2 / dS arithmetic of a pair is the critical path of the pipeline, the
// tensor-core work is not, so the row work is spread over twice the issue slots of the one-thread-per-row version.
#pragma once
#include <cstdio>
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
using tc::OBOX_B;
using tc::OutMaps;
using tc::obox_row;
using tc::obox_store;

constexpr int ROWS = 128;
constexpr int TILE_B = ROWS * 128;        // 16 KB
constexpr int STAGE_B = 2 * TILE_B;       // [Q | dO] and [K | V]
constexpr int NSTAGE = 3;
constexpr int GDEPTH = 1;               // a pair is published GDEPTH gather iterations after its copies were issued (2: the
                                        // publication then sits behind the NEXT stage wait + copy issue, measured slower)
constexpr int PD_B = 3 * 8192;            // block-diagonal P or dS: [data0 | zero | data1]
constexpr int BIAS_LD = 68;
constexpr int GATHER_WARP = 16;           // warps 0-15 rows (2 quads x 8), then NGW gather warps, then NMW MMA-issuing warps
constexpr int nthreads(int ngw, int nmw) { return 32 * (GATHER_WARP + ngw + nmw); }
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
  static constexpr int DPART = META + 2 * NSTAGE * ROWS * 4;     // [2 pairs in flight][2 quads][2 halves][128] fp32 partial rowsum(dO o O)
  static constexpr int BARS = DPART + 8 * ROWS * 4;
  static constexpr int TOTAL = BARS + 256;
};
static size_t bwd7_tc_smem() { return (size_t)Smem::TOTAL + 1024; }

// column sums of a [32 rows (lanes)] x [16 columns (v[0..15])] tile by a butterfly: at step `half` the lane whose bit `half`
// is set keeps the upper half of the live columns, so after four steps lane l holds the sum of column (l & 15) over the 16
// lanes that share bit 4 with it; one more exchange adds the other 16 rows.  16 shuffles per tensor.
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; j++) {
      const float mine = upper ? v[j + half] : v[j];
      const float send = upper ? v[j] : v[j + half];
      v[j] = mine + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}
// byte offset (within a quad's [P | dS] region) of output box idx = window * 3 + tensor: the data halves of the P tile and
// the first data half of the dS tile are free once the dQ / dK / dV GEMMs have completed (the shared zero regions are not
// touched)
__device__ __forceinline__ int obox_offset(int idx) {
  const int reg = idx >> 1;
  return (reg == 0 ? 0 : (reg == 1 ? 16384 : PD_B)) + (idx & 1) * OBOX_B;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

// PROF: per-role cycle accounting (ESVIT_ATTN_PROF=1): lane 0 of one warp per role charges the time since its previous tick
// to a named bucket; CTA (0, 0) prints the buckets at the end.  Development aid, not used by the product path.
#define WA_TICK(slot) do { if (PROF) { const long long t__ = clock64(); pacc[slot] += t__ - tlast; tlast = t__; } } while (0)

template <bool SHIFT, int NGW, int NMW, bool PROF = false>
__global__ void __launch_bounds__(nthreads(NGW, NMW), 1) window_attn_bwd7_tc_kernel(
    const bf16* __restrict__ qkv, const bf16* __restrict__ qkv_bias, const float* __restrict__ bexp,
    const bf16* __restrict__ out, const bf16* __restrict__ dout, const float* __restrict__ lse,
    bf16* __restrict__ dqkv, float* __restrict__ dbias_table, float* __restrict__ dqkv_bias, Geo g, float scale,
    int nwin_total, int dbg, const __grid_constant__ OutMaps om) {
  // dbg: 1 L2 prefetch of the O / dO rows by the gather warps (costs them ~400 cycles per prefetch instruction; off);  PROF builds only (results are then WRONG): 2 no dq/dk/dv stores, 4 no O / dO loads, 8 no copies
  constexpr int WS = 7, NT = 49, NB = 169;
  constexpr int NTHREADS = nthreads(NGW, NMW), MMA_WARP = GATHER_WARP + NGW;
  constexpr int RSTEP = 8 * NGW, RPT = (ROWS + RSTEP - 1) / RSTEP;   // gather: rows r = (t >> 2) + RSTEP kk < 128, kk < RPT
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = base + Smem::STAGES;
  uint8_t* pdbuf = base + Smem::PD;
  float* bias_s = reinterpret_cast<float*>(base + Smem::BIAS);
  float* bins = reinterpret_cast<float*>(base + Smem::BINS);
  int* tokb = reinterpret_cast<int*>(base + Smem::META);
  int* ridb = tokb + NSTAGE * ROWS;
  float* dpart = reinterpret_cast<float*>(base + Smem::DPART);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + Smem::BARS);
  uint64_t* full_in = bars;         // [3] count 64  (gather threads)
  uint64_t* empty_in = bars + 3;    // [3] count 1   (MMA commit after the second GEMM group)
  uint64_t* s_full = bars + 6;      // [2] count 1   (S and dP complete)
  uint64_t* pd_full = bars + 10;    // [2] count 8   (P and dS tiles written; S / dP consumed)
  uint64_t* g_full = bars + 12;     // [2] count 1   (dQ / dK / dV complete)
  uint64_t* g_free = bars + 14;     // [2] count 8   (row warps have loaded dQ / dK / dV: the accumulator columns are free)
  uint64_t* meta_full = bars + 16;  // [3] count 32 NGW (tok / rid of the stage written: published when the copies are ISSUED)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 20);

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
    for (int i = 0; i < NSTAGE; i++) { mbar_init(&full_in[i], 32 * NGW); mbar_init(&empty_in[i], 1); mbar_init(&meta_full[i], 32 * NGW); }
    for (int i = 0; i < 2; i++) {
      mbar_init(&s_full[i], 1); mbar_init(&pd_full[i], 8); mbar_init(&g_full[i], 1); mbar_init(&g_free[i], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;
  long long pacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = PROF ? clock64() : 0;

  if (warp >= GATHER_WARP && warp < MMA_WARP) {
    // ===================== gather warps: 64 threads, 4 lanes per 64-byte segment, 8 rows per thread =====================
    const int t = threadIdx.x - GATHER_WARP * 32;
    const int c16 = t & 3;
    uint4 bchunk[3];
#pragma unroll
    for (int part = 0; part < 3; part++)
      bchunk[part] = __ldg(reinterpret_cast<const uint4*>(qkv_bias + part * g.C + h * HD + c16 * 8));
    int iy[RPT], ix[RPT];   // slot geometry of this thread's rows (the same for every pair)
#pragma unroll
    for (int kk = 0; kk < RPT; kk++) {
      const int i = ((t >> 2) + RSTEP * kk) & 63;
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
      const int rr = (t >> 2) + RSTEP * kk;
      const int w = (rr >> 6) & 1;
      tk = -2; rd = 0;
      if (rr < ROWS && pg.ok[w] && iy[kk] < WS) {
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
    for (int it = 0; it < n_items + GDEPTH; it++) {
      // FIRST publish the pair issued GDEPTH iterations ago (its copies have landed), THEN wait for a free stage: the stage
      // this iteration needs is released by GEMMs that themselves wait for that publication (circular otherwise)
      WA_TICK(0);
      if (it >= GDEPTH) {
        cp_async_wait<GDEPTH - 1>();
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        mbar_arrive(&full_in[(it - GDEPTH) % NSTAGE]);
      }
      WA_TICK(1);
      if (it < n_items) {
        const int st_i = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        const int pair = blockIdx.y + it * gridDim.y;
        if (it + NSTAGE < n_items && (dbg & 1)) {  // L2 prefetch of the pair three ahead (q / k / v segments, dO and O rows)
          const PairGeo pf = pair_geo(wpf, pair + NSTAGE * (int)gridDim.y);
#pragma unroll
          for (int kk = 0; kk < RPT; kk++) {
            int tk, rd;
            slot(pf, kk, tk, rd);
            if (tk >= 0) {   // the rows the row threads read straight from global memory; dbg & 64: the q / k / v segments too
              if (c16 == 0) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(dout + (long long)tk * g.C + h * HD));
              else if (c16 == 1) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(out + (long long)tk * g.C + h * HD));
              else if (dbg & 64) {
                asm volatile("prefetch.global.L2 [%0];\n" ::"l"(qkv + (long long)tk * 3 * g.C + (c16 - 2) * g.C + h * HD));
                if (c16 == 3) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(qkv + (long long)tk * 3 * g.C + 2 * g.C + h * HD));
              }
            }
          }
        }
        const PairGeo pg = pair_geo(wcur, pair);
        advance(wcur, wstep);
        advance(wpf, wstep);
        WA_TICK(2);
        mbar_wait(&empty_in[st_i], ph ^ 1);
        WA_TICK(3);
        uint8_t* t1 = stages + st_i * STAGE_B;   // [Q | dO]
        uint8_t* t2 = t1 + TILE_B;               // [K | V]
#pragma unroll
        for (int kk = 0; kk < RPT; kk++) {
          const int r = (t >> 2) + RSTEP * kk;
          if (ROWS % RSTEP != 0 && r >= ROWS) break;
          int tk, rd;
          slot(pg, kk, tk, rd);
          const int sw = r & 7;
          uint8_t* dq = t1 + r * 128 + ((c16 ^ sw) * 16);
          uint8_t* dd = t1 + r * 128 + (((4 + c16) ^ sw) * 16);
          uint8_t* dk = t2 + r * 128 + ((c16 ^ sw) * 16);
          uint8_t* dv = t2 + r * 128 + (((4 + c16) ^ sw) * 16);
          if (PROF && (dbg & 8) && it > 0) tk = -2;
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
        mbar_arrive(&meta_full[st_i]);   // the row threads fetch their per-row global data one pair ahead of the tiles
      }
      cp_async_commit();
    }
    WA_TICK(0);
    if (PROF && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == GATHER_WARP * 32)
      printf("bwd7_tc gather: items %d | copy-issue %lld  cp.async.wait %lld  prefetch+geometry %lld  wait empty_in %lld\n", n_items,
             pacc[0], pacc[1], pacc[2], pacc[3]);
  } else if (warp >= MMA_WARP) {
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
      auto issue_s = [&](int it) {   // S(it), dP(it)
        const int b = it & 1, st_i = it % NSTAGE;
        WA_TICK(0);
        mbar_wait(&full_in[st_i], (it / NSTAGE) & 1);
        WA_TICK(1);
        mbar_wait(&g_free[b], ((it >> 1) & 1) ^ 1);   // the quad has drained dQ / dK / dV of its previous pair
        WA_TICK(2);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint64_t so = (uint64_t)(st_i * (STAGE_B >> 4));
#pragma unroll
        for (int k = 0; k < 2; k++) umma(tmem_base + b * BUF_COLS, t1k + so + 2 * k, t2k + so + 2 * k, id_s, k);                 // Q K^T
#pragma unroll
        for (int k = 0; k < 2; k++) umma(tmem_base + b * BUF_COLS + 128, t1k + so + 4 + 2 * k, t2k + so + 4 + 2 * k, id_s, k);   // dO V^T
        umma_commit(&s_full[b]);
      };
      auto issue_g = [&](int j) {    // dQ, dK, dV of pair j
        const int b = j & 1, st_j = j % NSTAGE;
        const uint32_t ph = (j >> 1) & 1;
        WA_TICK(0);
        mbar_wait(&pd_full[b], ph);                   // implies the quad has loaded S / dP: their columns are free
        WA_TICK(3);
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
      };
      if (NMW == 1) {
        // one issuing thread: S / dP of pair it, then dQ / dK / dV of pair it - 1 (the other quad's)
        for (int it = 0; it <= n_items; it++) {
          if (it < n_items) issue_s(it);
          if (it > 0) issue_g(it - 1);
        }
      } else {
        // one issuing thread PER QUAD: a quad that is late (its accumulators not drained, its P / dS not written) does not
        // hold up the other quad's GEMMs behind it in a single thread's program order
        for (int it = warp - MMA_WARP; it < n_items; it += 2) {
          issue_s(it);
          issue_g(it);
        }
      }
      WA_TICK(0);
      if (PROF && blockIdx.x == 0 && blockIdx.y == 0 && warp == MMA_WARP)
        printf("bwd7_tc mma: issue %lld  wait full_in %lld  wait g_free %lld  wait pd_full %lld\n", pacc[0], pacc[1], pacc[2], pacc[3]);
    }
  } else {
    // ===================== row warps: two quads of EIGHT warps, TWO threads per row =====================
    // warp = quad * 8 + half * 4 + lq: lq = the TMEM lane quarter (warp id % 4), half = which 32 of the row's 64 key columns
    // (phase 1) / which 16 of the 32 channels (phase 2) the thread owns.
    const int quad = warp >> 3, hh = (warp >> 2) & 1, lq = warp & 3;
    const int r = lq * 32 + lane;
    const int w = r >> 6, i = r & 63;
    const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + quad * BUF_COLS;
    const float c = scale * LOG2E;
    const float* brow = bias_s + i * BIAS_LD + hh * 32;
    uint8_t* obase = pdbuf + quad * 2 * PD_B;                      // the quad's [P | dS] region, re-used for the output boxes
    uint8_t* prow = obase + w * 8192 + r * 128;                     // P row; dS row at + PD_B
    const int siy = i / WS, six = i - siy * WS;
    const bool issuer = lane == 0 && (warp & 7) < 6;   // issues the bulk tensor stores of one output box per pair
    float dsacc[32];
#pragma unroll
    for (int j = 0; j < 32; j++) dsacc[j] = 0.f;
    float csum[3] = {0.f, 0.f, 0.f};  // per-lane column sums of dQ, dK, dV (channel hh * 16 + (lane & 15))
    // what a pair needs from global memory per row, fetched one pair of this quad AHEAD (while the quad waits for the
    // dQ / dK / dV GEMMs of the current pair; the L2 round trip was 20 % of the quad's time when it sat at the top of the
    // iteration): D = rowsum(dO o O) -- this thread's 16 channels, left in shared memory for the exchange with the row's
    // other thread -- and the row's LSE.  Padded query rows (tok == -1) have dO = 0, hence dS = 0 and no dV contribution:
    // P = exp2(-inf) = 0 serves them too (and keeps an undefined saved LSE of a skipped all-padding tile out of the arithmetic)
    auto prep = [&](int it2, int& rid_o, float& l2_o) {
      const int st2 = it2 % NSTAGE;
      WA_TICK(7);
      mbar_wait(&meta_full[st2], (it2 / NSTAGE) & 1);
      WA_TICK(1);
      const int tok = tokb[st2 * ROWS + r];
      rid_o = SHIFT ? ridb[st2 * ROWS + r] : 0;
      float Dr = 0.f, l2 = INFINITY;
      if (tok >= 0 && !(PROF && (dbg & 4))) {
        const int pair2 = blockIdx.y + it2 * gridDim.y;
        const uint4* po = reinterpret_cast<const uint4*>(out + (long long)tok * g.C + h * HD + hh * 16);
        const uint4* pd = reinterpret_cast<const uint4*>(dout + (long long)tok * g.C + h * HD + hh * 16);
        const uint4 uo0 = __ldg(po), uo1 = __ldg(po + 1), ud0 = __ldg(pd), ud1 = __ldg(pd + 1);
        l2 = __ldg(lse + ((long long)(2 * pair2 + w) * g.nH + h) * NT + i) * LOG2E;
        float fo[8], fd[8];
        unpack8(*reinterpret_cast<const bf16x8*>(&uo0), fo);
        unpack8(*reinterpret_cast<const bf16x8*>(&ud0), fd);
#pragma unroll
        for (int e = 0; e < 8; e++) Dr = fmaf(fo[e], fd[e], Dr);
        unpack8(*reinterpret_cast<const bf16x8*>(&uo1), fo);
        unpack8(*reinterpret_cast<const bf16x8*>(&ud1), fd);
#pragma unroll
        for (int e = 0; e < 8; e++) Dr = fmaf(fo[e], fd[e], Dr);
      }
      dpart[((((it2 >> 1) & 1) * 2 + quad) * 2 + hh) * ROWS + r] = Dr;
      l2_o = l2;
      WA_TICK(2);
    };
    int rid_r = 0;
    float l2 = INFINITY;
    if (quad < n_items) prep(quad, rid_r, l2);
    for (int it = quad; it < n_items; it += 2) {
      const int st_i = it % NSTAGE;
      const uint32_t ph = (it >> 1) & 1;
      const int pair = blockIdx.y + it * gridDim.y;
      const float* dbuf = dpart + (((it >> 1) & 1) * 2 + quad) * 2 * ROWS + r;
      mbar_wait(&s_full[quad], ph);
      WA_TICK(3);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      // (the quad's bulk stores of the previous pair have finished READING the P / dS regions: awaited by the issuing
      // threads before this barrier)
      if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
      asm volatile("bar.sync %0, 256;\n" ::"r"(quad + 1) : "memory");
      const float Dr = dbuf[0] + dbuf[ROWS];
      WA_TICK(4);
      // the quad's previous dQ / dK / dV GEMMs have completed (g_full awaited below in the previous iteration), so the
      // P / dS tiles may be overwritten
#pragma unroll
      for (int q16 = 0; q16 < 2; q16++) {
        uint32_t sv[16], dv[16];
        tmem_ld16(taddr + w * 64 + hh * 32 + q16 * 16, sv);
        tmem_ld16(taddr + 128 + w * 64 + hh * 32 + q16 * 16, dv);
        tmem_ld_wait();
#pragma unroll
        for (int c2 = 0; c2 < 2; c2++) {
          const int ch = q16 * 2 + c2;
          float p8[8], d8[8];
          const float4 b0 = *reinterpret_cast<const float4*>(brow + ch * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(brow + ch * 8 + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          int rc[8];
          if (SHIFT) {
            const int4 r0 = *reinterpret_cast<const int4*>(ridb + st_i * ROWS + w * 64 + hh * 32 + ch * 8);
            const int4 r1 = *reinterpret_cast<const int4*>(ridb + st_i * ROWS + w * 64 + hh * 32 + ch * 8 + 4);
            rc[0] = r0.x; rc[1] = r0.y; rc[2] = r0.z; rc[3] = r0.w; rc[4] = r1.x; rc[5] = r1.y; rc[6] = r1.z; rc[7] = r1.w;
          }
#pragma unroll
          for (int e = 0; e < 8; e++) {
            const int jl = ch * 8 + e;          // column within this thread's half
            float x = fmaf(__uint_as_float(sv[c2 * 8 + e]), c, bb[e]) - l2;
            if (SHIFT) { if (hh * 32 + jl < NT && rc[e] != rid_r) x += -100.f * LOG2E; }
            const float pj = ex2(x);
            const float dsj = pj * (__uint_as_float(dv[c2 * 8 + e]) - Dr);
            p8[e] = pj;
            d8[e] = dsj;
            dsacc[jl] += dsj;
          }
          const int sw = ((hh * 4 + ch) ^ (r & 7)) * 16;
          *reinterpret_cast<bf16x8*>(prow + sw) = pack8(p8);
          *reinterpret_cast<bf16x8*>(prow + PD_B + sw) = pack8(d8);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&pd_full[quad]);
      WA_TICK(5);
      int rid_n = 0;
      float l2_n = INFINITY;
      if (it + 2 < n_items) prep(it + 2, rid_n, l2_n);
      // phase 2: dQ / dK / dV of this pair -> bf16 -> token order; column sums -> qkv-bias gradient.  This thread: channels
      // hh * 16 .. + 15 of its row in each tensor: dQ at +0, dK at +32, dV at +96 (columns 32..63 of [. | dV])
      mbar_wait(&g_full[quad], ph);
      WA_TICK(6);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      uint32_t gr[3][16];
      tmem_ld16(taddr + hh * 16, gr[0]);
      tmem_ld16(taddr + 32 + hh * 16, gr[1]);
      tmem_ld16(taddr + 96 + hh * 16, gr[2]);
      tmem_ld_wait();
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&g_free[quad]);   // the accumulator columns may be overwritten
      WA_TICK(8);
      // row of this slot inside its window's staged box.  A window that wraps around the image (last window row / column of
      // a shifted block) is staged as 2 or 4 dense sub-boxes, one per wrapped copy: [n1 | shift] columns x [n1 | shift] rows
      const int orow = SHIFT ? obox_row(siy, six, (rid_r & 16) != 0, (rid_r & 32) != 0, g.shift) : i;
#pragma unroll
      for (int part = 0; part < 3; part++) {
        float f[16];
        const float sc = part == 2 ? 1.f : scale;
#pragma unroll
        for (int e = 0; e < 16; e++) f[e] = __uint_as_float(gr[part][e]) * sc;
        if (i < NT) {   // this row's 32 bytes of box (w, part): chunks 2 hh, 2 hh + 1 of its 64-byte row
          uint8_t* box = obase + obox_offset(w * 3 + part) + orow * 64;
          const int sw = (orow >> 1) & 3;
          *reinterpret_cast<bf16x8*>(box + (((2 * hh) ^ sw) * 16)) = pack8(f);
          *reinterpret_cast<bf16x8*>(box + (((2 * hh + 1) ^ sw) * 16)) = pack8(f + 8);
        }
        // qkv-bias gradient: column sums over ALL slots of the window (padded ones included; missing rows are 0)
        csum[part] += warp_colsum16(f, lane);
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      WA_TICK(9);
      asm volatile("bar.sync %0, 256;\n" ::"r"(quad + 1) : "memory");
      WA_TICK(10);
      // six issuing threads per quad (lane 0 of its first six warps), one per (window, tensor) box
      if (issuer && !(PROF && (dbg & 2))) {
        const int ww = (warp & 7) >= 3, part = (warp & 7) - 3 * ww;
        const int win = 2 * pair + ww;
        if (win < nwin_total) obox_store<SHIFT>(om, smem_u32(obase) + obox_offset(ww * 3 + part), part * g.C + h * HD, win, g);
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
      rid_r = rid_n;
      l2 = l2_n;
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    WA_TICK(7);
    if (PROF && blockIdx.x == 0 && blockIdx.y == 0 && ((threadIdx.x & 255) == 0 || (threadIdx.x & 255) == 160))
      printf("bwd7_tc row quad %d warp %d: wait meta %lld  O/dO/lse loads %lld  wait s_full %lld  D exchange %lld  phase 1 %lld  "
             "wait g_full %lld | phase 2: tmem ld %lld  stage+colsum %lld  bar %lld  store issue+loop %lld\n", quad, warp & 7, pacc[1],
             pacc[2], pacc[3], pacc[4], pacc[5], pacc[6], pacc[8], pacc[9], pacc[10], pacc[7]);
    // flush: rel-pos-bias gradient through the CTA's shared-memory bins, qkv-bias gradient straight to global memory
    if (i < NT) {
#pragma unroll
      for (int j = 0; j < 32; j++)
        if (hh * 32 + j < NT) atomicAdd(&bins[bias_index<WS>(i, hh * 32 + j)], dsacc[j]);
    }
    if (lane < 16) {
#pragma unroll
      for (int part = 0; part < 3; part++) atomicAdd(&dqkv_bias[part * g.C + h * HD + hh * 16 + lane], csum[part]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  for (int i = threadIdx.x; i < NB; i += NTHREADS) atomicAdd(&dbias_table[i * g.nH + h], bins[i]);
  if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS));
}

}  // namespace tcb
}  // namespace wa
