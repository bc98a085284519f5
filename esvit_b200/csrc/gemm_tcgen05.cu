// Blackwell-native bf16 GEMM with a fused epilogue:  D[M,N] = act( A[M,K] . W[N,K]^T + bias[N] )   (nn.Linear layout)
//
//   * operands staged by TMA (cp.async.bulk.tensor.2d, 128B-swizzled K-major tiles) into a 4-stage mbarrier ring,
//   * tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) issued by ONE elected thread, accumulator in TMEM
//     (128 lanes x 128 columns, double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1),
//   * 16 epilogue warps (the bias + exact-GELU math is latency-bound, it needs the warps) read the accumulator with tcgen05.ld, add the bias, optionally apply exact GELU (also keeping the
//     pre-activation that the backward needs), stage bf16 tiles in 128B-swizzled shared memory and hand them to TMA
//     stores (cp.async.bulk.tensor ... global.shared::cta), which also clip the M / N tails.
// Persistent: grid = #SMs, static round-robin over (m_tile, n_tile) with n fastest so the 128-row A tile is re-read
// from L2 by neighbouring CTAs.  K/M/N tails rely on TMA zero-fill and an epilogue column/row mask.
//
// This is the fc1 (+bias +GELU) GEMM of the Swin MLP (models/swin_transformer.py:31-33) and the generic "Linear" of
// the block; the descriptor encodings follow the sm_100 UMMA definitions (InstrDescriptor / SmemDescriptor).
#include <cuda.h>

#include "common.cuh"

namespace tg {

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3;
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NUM_EPI_WARPS = 16, NUM_THREADS = 32 * (3 + NUM_EPI_WARPS);  // warp0 TMA, warp1 MMA, warps 2-17 epilogue, warp18 multiplier TMA
constexpr int TMEM_COLS = 2 * BN;                                          // double-buffered accumulator
constexpr int OUT_TILE_BYTES = BM * BN * 2;                                // bf16 staging tile of the TMA store
constexpr int OUT_HALF_BYTES = BM * 64 * 2;                                // one 64-column (128-byte) swizzled box

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// bounded spin: a mis-programmed pipeline traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 28); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;\n" ::"n"(NUM_EPI_WARPS * 32) : "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes (64 bf16), 8-row swizzle atoms 1024 B apart.
// SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;             // leading byte offset (unused for swizzled K-major), canonical value 1
  d |= (uint64_t)(1024 >> 4) << 32;   // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}
// InstrDescriptor (kind::f16): c_format F32=1 [4,6) | a_format BF16=1 [7,10) | b_format BF16=1 [10,13) |
// a_major K=0 [15] | b_major K=0 [16] | n_dim=N>>3 [17,23) | m_dim=M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// exact-erf GELU and its derivative from one exp (A&S 7.1.26 erf, same polynomial as elementwise.cu):
// gelu(x) = x*Phi(x), gelu'(x) = Phi(x) + x*phi(x)
__device__ __forceinline__ void gelu_and_grad(float x, float& y, float& dy) {
  const float e = __expf(-0.5f * x * x);
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float cdf = 0.5f * (1.f + copysignf(1.f - p * t * e, x));
  y = x * cdf;
  dy = fmaf(x * 0.3989422804014327f, e, cdf);
}

struct Params {
  bf16* out;        // [M, N] act(A W^T + bias)
  bf16* pre;        // [M, N] d act/dx at the pre-activation (gelu'(A W^T + bias)), written when act == 1 and pre != nullptr
  const float* bias;  // [N] or nullptr
  float* colsum;      // act == 2: [gridDim.x][N] per-CTA partial column sums of out (zeroed by the launcher)
  int M, N, K, act;   // act: 0 = identity, 1 = exact GELU, 2 = out = (A W^T) * mult[M,N] (mult arrives through map_pre)
};
constexpr int ACT_MUL = 2;

__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_bias_act_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                       const __grid_constant__ CUtensorMap map_b,
                                                                       const __grid_constant__ CUtensorMap map_out,
                                                                       const __grid_constant__ CUtensorMap map_pre,
                                                                       const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // [STAGES][A | B], every tile 1024-byte aligned in the SHARED address space (swizzle-128B requirement)
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // store staging, double-buffered over tiles: [2 buffers][out | pre][2 halves][128 rows][128 B] (128B-swizzled)
  uint8_t* stage_base = tiles + STAGES * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(stage_base + 4 * OUT_TILE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint64_t* mult_full = tmem_empty + 2;   // [2] act == 2: multiplier tile landed in the `pre` staging slot
  uint64_t* mult_empty = mult_full + 2;   // [2] ... and has been consumed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(mult_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN, k_blocks = (p.K + BK - 1) / BK;
  const int num_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; i++) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], NUM_EPI_WARPS * 32); }
    for (int i = 0; i < 2; i++) { mbar_init(&mult_full[i], 1); mbar_init(&mult_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 2) {  // one warp allocates TMEM (and frees it at the end)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < k_blocks; kb++) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = tiles + stage * STAGE_BYTES;
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          tma_load_2d(sa, &map_a, &full[stage], kb * BK, m0);
          tma_load_2d(sa + A_BYTES, &map_b, &full[stage], kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 2 + NUM_EPI_WARPS) {
    // ===================== multiplier stream (act == 2) =====================
    // its own warp: waiting for a staging slot must not hold back the operand prefetch.  The tile's multiplier goes
    // straight into the staging slot the epilogue reads it from.
    if (p.act == ACT_MUL && elect_one()) {
      int local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, local++) {
        const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
        const int as = local & 1;
        mbar_wait(&mult_empty[as], ((local >> 1) & 1) ^ 1);
        uint8_t* sm = stage_base + as * 2 * OUT_TILE_BYTES + OUT_TILE_BYTES;
        const bool two = n0 + 64 < p.N;
        mbar_expect_tx(&mult_full[as], two ? OUT_TILE_BYTES : OUT_HALF_BYTES);
        tma_load_2d(sm, &map_pre, &mult_full[as], n0, m0);
        if (two) tma_load_2d(sm + OUT_HALF_BYTES, &map_pre, &mult_full[as], n0 + 64, m0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected thread) =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, local++) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);  // epilogue has drained this accumulator stage
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; kb++) {
          mbar_wait(&full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint32_t sa = smem_u32(tiles + stage * STAGE_BYTES);
          const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[as]);   // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue warps: TMEM -> registers -> (+bias, GELU) -> swizzled smem -> TMA store ==========
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int colq = (warp - 2) >> 2;          // 32-column slice of the 128-column tile handled by this warp
    const bool store_pre = p.act && p.pre;
    const int r = q * 32 + lane;               // row of the tile owned by this thread
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, local++) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
      uint8_t* stage_out = stage_base + as * 2 * OUT_TILE_BYTES;
      uint8_t* stage_pre = stage_out + OUT_TILE_BYTES;
      uint8_t* so = stage_out + (colq >> 1) * OUT_HALF_BYTES + r * 128;
      uint8_t* sp = stage_pre + (colq >> 1) * OUT_HALF_BYTES + r * 128;
      mbar_wait(&tmem_full[as], aphase);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + colq * 32), v);
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      mbar_arrive(&tmem_empty[as]);  // accumulator drained into registers: the MMA warp may start the next-but-one tile
      // the TMA stores issued two tiles ago (same staging buffer) must have finished READING it
      if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
      epi_bar();
      if (p.act == ACT_MUL) {
        // out = acc * mult (the fc2 input-gradient GEMM fused with the GELU backward: mult = gelu'(pre-activation)),
        // column sums of out = the fc1 bias gradient
        mbar_wait(&mult_full[as], aphase);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int sw = (((colq & 1) * 4 + j) ^ (r & 7)) * 16;
          float fm[8], g[8];
          unpack8(*reinterpret_cast<const bf16x8*>(sp + sw), fm);
#pragma unroll
          for (int t = 0; t < 8; t++) g[t] = __uint_as_float(v[j * 8 + t]) * fm[t];
          *reinterpret_cast<bf16x8*>(so + sw) = pack8(g);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        epi_bar();
        if (threadIdx.x == 64) {
          mbar_arrive(&mult_empty[as]);  // every epilogue thread is past its multiplier reads
          tma_store_2d(&map_out, stage_out, n0, m0);
          if (n0 + 64 < p.N) tma_store_2d(&map_out, stage_out + OUT_HALF_BYTES, n0 + 64, m0);
          asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
        {  // column sums over the staged bf16 tile: thread = (column pair, 16-row group); rows >= M hold zeros
          const int te = threadIdx.x - 64, cp = te & 63, rg = te >> 6;
          const uint8_t* base = stage_out + (cp >> 5) * OUT_HALF_BYTES + (cp & 3) * 4;
          const int chunk = (cp & 31) >> 2;
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const int row = rg * 16 + i;
            const uint32_t w2 = *reinterpret_cast<const uint32_t*>(base + row * 128 + ((chunk ^ (row & 7)) * 16));
            s0 += __uint_as_float(w2 << 16);
            s1 += __uint_as_float(w2 & 0xffff0000u);
          }
          // into this CTA's PRIVATE row: all CTAs adding into one [N] vector serialise on its few cache lines in the
          // L2 atomic units (N = 384 is 12 lines: measured 2x the whole kernel)
          float* dst = p.colsum + (long long)blockIdx.x * p.N;
          const int col = n0 + cp * 2;
          if (col < p.N) atomicAdd(dst + col, s0);
          if (col + 1 < p.N) atomicAdd(dst + col + 1, s1);
        }
        continue;
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {  // 4 chunks of 8 columns = 16 B
        const int cbase = n0 + colq * 32 + j * 8;
        float f[8], g[8];  // g = act(x); f = what the `pre` slot receives: d act / dx (act = GELU), else unused
#pragma unroll
        for (int t = 0; t < 8; t++) {
          const float x = __uint_as_float(v[j * 8 + t]) + ((p.bias && cbase + t < p.N) ? p.bias[cbase + t] : 0.f);
          if (p.act) {
            gelu_and_grad(x, g[t], f[t]);
          } else {
            g[t] = x;
            f[t] = x;
          }
        }
        const int sw = (((colq & 1) * 4 + j) ^ (r & 7)) * 16;  // 128B swizzle: 16-byte chunk index XOR (row mod 8)
        *reinterpret_cast<bf16x8*>(so + sw) = pack8(g);
        if (store_pre) *reinterpret_cast<bf16x8*>(sp + sw) = pack8(f);
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy smem writes -> visible to TMA
      epi_bar();
      if (threadIdx.x == 64) {
        tma_store_2d(&map_out, stage_out, n0, m0);
        if (n0 + 64 < p.N) tma_store_2d(&map_out, stage_out + OUT_HALF_BYTES, n0 + 64, m0);
        if (store_pre) {
          tma_store_2d(&map_pre, stage_pre, n0, m0);
          if (n0 + 64 < p.N) tma_store_2d(&map_pre, stage_pre + OUT_HALF_BYTES, n0 + 64, m0);
        }
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
    }
    if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");  // stores complete before exit
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)ptr;
  }
  return fn;
}

// row-major [rows, cols] bf16 matrix, box = [box_rows, 64 cols], 128B swizzle, zero fill out of bounds
static bool make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, int box_rows) {
  EncodeFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// colsum[n] += sum over the CTAs' private rows
__global__ void __launch_bounds__(256) colsum_fold_kernel(const float* __restrict__ ws, int rows, int N, float* __restrict__ colsum) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int r = 0; r < rows; r++) s += ws[(long long)r * N + n];
  colsum[n] += s;
}

constexpr int MAX_GRID = 160;  // bound on the persistent grid = rows of the column-sum workspace

static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const CUtensorMap& mp,
                  const Params& p, void* stream, int* grid_out = nullptr) {
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 4 * OUT_TILE_BYTES + (2 * STAGES + 8) * sizeof(uint64_t) + 16 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bias_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  int grid = esvit_num_sms();
  if (grid > tiles) grid = tiles;
  if (grid > MAX_GRID) grid = MAX_GRID;
  if (p.act == ACT_MUL) {
    cudaError_t e = cudaMemsetAsync(p.colsum, 0, (size_t)grid * p.N * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
  }
  if (grid_out) *grid_out = grid;
  gemm_bias_act_kernel<<<grid, NUM_THREADS, smem, (cudaStream_t)stream>>>(ma, mb, mo, mp, p);
  ESVIT_LAUNCH_CHECK();
}

}  // namespace tg

// out[M,N] (bf16) = act(a[M,K] @ w[N,K]^T + bias[N]);  act: 0 identity, 1 exact GELU (then `pre`, if not NULL, receives
// gelu'(pre-activation), the local derivative the backward multiplies with).  a, w bf16 row-major with K contiguous (K % 8 == 0, N % 8 == 0), bias fp32 or NULL.
ESVIT_API int esvit_gemm_bias_act(const void* a, const void* w, const float* bias, void* out, void* pre, long long M,
                                  int N, int K, int act, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (N % 8) || (K % 8) || M > 0x7fffffffLL) return ESVIT_ERR_BAD_ARG;
  CUtensorMap ma, mb, mo, mp;
  if (!tg::make_map(&ma, a, M, K, tg::BM) || !tg::make_map(&mb, w, N, K, tg::BN)) return ESVIT_ERR_BAD_ARG;
  if (!tg::make_map(&mo, out, M, N, tg::BM) || !tg::make_map(&mp, (act && pre) ? pre : out, M, N, tg::BM))
    return ESVIT_ERR_BAD_ARG;
  tg::Params p;
  p.out = (bf16*)out; p.pre = (bf16*)pre; p.bias = bias; p.colsum = nullptr; p.M = (int)M; p.N = N; p.K = K; p.act = act;
  return tg::launch(ma, mb, mo, mp, p, stream);
}

// out[M,N] (bf16) = (a[M,K] @ w[N,K]^T) * mult[M,N];  colsum[N] (fp32) += column sums of out.
// The input-gradient GEMM of fc2 fused with the GELU backward of fc1: a = dy, w = W2^T, mult = gelu'(pre-activation) as
// saved by esvit_gemm_bias_act(act = 1), out = d(pre-activation), colsum = fc1 bias gradient (caller zero-fills);
// ws fp32 [160 * N]: scratch for the per-CTA partial column sums.
ESVIT_API int esvit_gemm_mul_colsum(const void* a, const void* w, const void* mult, void* out, float* colsum, float* ws,
                                    long long M, int N, int K, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (N % 8) || (K % 8) || M > 0x7fffffffLL || !mult || !colsum || !ws) return ESVIT_ERR_BAD_ARG;
  CUtensorMap ma, mb, mo, mp;
  if (!tg::make_map(&ma, a, M, K, tg::BM) || !tg::make_map(&mb, w, N, K, tg::BN)) return ESVIT_ERR_BAD_ARG;
  if (!tg::make_map(&mo, out, M, N, tg::BM) || !tg::make_map(&mp, mult, M, N, tg::BM)) return ESVIT_ERR_BAD_ARG;
  tg::Params p;
  p.out = (bf16*)out; p.pre = nullptr; p.bias = nullptr; p.colsum = ws; p.M = (int)M; p.N = N; p.K = K;
  p.act = tg::ACT_MUL;
  int rows = 0;  // rows of ws in use = CTAs launched
  const int rc = tg::launch(ma, mb, mo, mp, p, stream, &rows);
  if (rc != 0) return rc;
  tg::colsum_fold_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ws, rows, N, colsum);
  ESVIT_LAUNCH_CHECK();
}
