// Blackwell-native bf16 GEMM family (second generation): every Linear of the training step - forward, input gradient
// and weight gradient - on tcgen05 tensor cores with TMA-staged operands and TMEM accumulators.
//
//   D[M,N] = epi( opA(A) . opB(B) )        fp32 accumulation in TMEM
//
//   * CTA PAIRS (cta_group::2, template CG = 2): two CTAs of a cluster work on one 256 x BN tile; each stages its own
//     128 rows of A and HALF of B (the pair reads B once), the leader CTA issues tcgen05.mma.cta_group::2 with M = 256,
//     each CTA drains its own 128 accumulator rows.  CG = 1 is the single-CTA form of the same kernel.
//   * operands K-major (nn.Linear forward: A[M,K], W[N,K]) or MN-major (a_mn / b_mn): the SAME row-major matrices
//     read "transposed" by TMA boxes of [64 contiguous MN elements x 64 K rows] and UMMA MN-major descriptors, so
//     the input-gradient GEMM dX = dY . W reads W[N,K] as it lies and the weight-gradient GEMM dW = dY^T . X reads
//     dY[T,N] and X[T,K] as they lie: no transposed copies exist anywhere.
//   * the MMA N extent is a run-time field of the instruction descriptor: the last N tile issues only the columns
//     that exist (N = 96 costs 96 columns of tensor time, not 128 / 256).
//   * epilogues (template EPI): bias -> bf16 | bias + exact GELU (+ gelu' for the backward) | accumulator * TMA-loaded
//     multiplier + column sums (fc2 dgrad fused with GELU' and the fc1 bias gradient) | fp32 split-K partials for the
//     weight gradients (deterministic: partial tiles + fold kernel, no atomics).
//   * 16 epilogue warps = FOUR INDEPENDENT quads; a quad owns one 128-row x 64-column slab at a time (tcgen05.ld -> math ->
//     128B-swizzled smem slab -> its own TMA store stream, its own named barrier), slabs are dealt round-robin over the
//     quads, so four slabs of up to two tiles are in flight per SM and no barrier ever spans more than 128 threads.
// Persistent: one CTA per SM, static round-robin over (split, m tile, n tile) work items.
#include <cuda.h>

#include "common.cuh"

namespace tg2 {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 16, NUM_THREADS = 32 * (3 + NUM_EPI_WARPS);  // warp0 TMA, warp1 MMA, warps 2-17 epilogue, warp18 multiplier TMA
constexpr int A_BYTES = BM * BK * 2;             // 16 KB: 128 rows x 64 K (either major)
constexpr int SLAB_BYTES = BM * 64 * 2;          // 16 KB: 128 rows x 64 output columns, one 128B-swizzled TMA box
constexpr int SMEM_LIMIT = 232448;               // 227 KB opt-in dynamic shared memory per CTA
enum { EPI_BIAS = 0, EPI_GELU = 1, EPI_MUL = 2, EPI_F32 = 3 };

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
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// bounded spin: a mis-programmed pipeline traps after ~2 s of wall time instead of hanging the GPU
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
template <bool CLUSTER>
__device__ __forceinline__ void mbar_wait_impl(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (uint32_t it = 0;; ++it) {
    if constexpr (CLUSTER) {  // acquire at cluster scope: completed by the peer CTA / the async proxy of the pair
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(a), "r"(parity)
          : "memory");
    } else {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(a), "r"(parity)
          : "memory");
    }
    if (done) return;
    if ((it & 1023u) == 1023u) {
      const uint64_t t = global_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 2000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_impl<false>(bar, parity); }
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) { mbar_wait_impl<true>(bar, parity); }
template <int CG>
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_addr, int c0, int c1) {
  if constexpr (CG == 2) {  // both CTAs of the pair issue their own loads; completion bytes go to the LEADER's barrier
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_dst),
        "l"(map), "r"(bar_addr), "r"(c0), "r"(c1)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_dst),
        "l"(map), "r"(bar_addr), "r"(c0), "r"(c1)
        : "memory");
  }
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// named barrier of one epilogue quad (4 warps): ids 1..4
__device__ __forceinline__ void quad_bar(int g) { asm volatile("bar.sync %0, 128;\n" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}

// Shared-memory matrix descriptors, 128B swizzle (SmemDescriptor of the sm_100 UMMA definitions):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
//   K-major tile : rows of 128 B (64 bf16 along K), 8-row swizzle atoms 1024 B apart (SBO); LBO unused (canonical 1);
//                  one MMA (K = 16) advances the start address by 32 B.
//   MN-major tile: TMA boxes of [64 K rows][64 MN elements = 128 B]: 8-K-row atoms 1024 B apart (SBO), the next 64 MN
//                  elements (next box) 8192 B further (LBO); one MMA (K = 16 rows) advances the start by 2048 B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, bool mn_major) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(mn_major ? (8192 >> 4) : 1) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// InstrDescriptor (kind::f16): c_format F32=1 [4,6) | a_format BF16=1 [7,10) | b_format BF16=1 [10,13) |
// a_major [15] | b_major [16] (0 = K, 1 = MN) | n_dim=N>>3 [17,23) | m_dim=M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if constexpr (CG == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  }
}
// MMA completion -> mbarrier at this smem offset (CG = 2: in BOTH CTAs of the pair)
template <int CG>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  if constexpr (CG == 2) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
  }
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

// GELU in the epilogue: Phi(x) = 0.5 + 0.5 tanh(x (a + b x^2)) with ONE MUFU (tanh.approx) and 5 FMA-pipe instructions;
// |gelu - x Phi_erf(x)| < 1e-3 absolute at |x| ~ 4 and < 3e-4 for |x| < 3 - below the bf16 resolution of the stored value
// (0.4 % relative).  The stored derivative is the derivative of THIS function (0 MUFU, 5 more instructions), so forward
// and backward are consistent.  The epilogue is instruction-bound, not HBM-bound: the erf form (2 MUFU + 14 ALU) and a
// degree-15 minimax polynomial (19 ALU) both measured slower than the HBM write stream they feed.
constexpr float GELU_A = 0.7978845608028654f, GELU_B = 0.7978845608028654f * 0.044715f;
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
// Two elements per instruction (fma.rn.f32x2 / mul.rn.f32x2, sm_100): the fp32 FMA pipe issues one instruction per two cycles
// and SMSP, and at 11 of them per element the GELU epilogue, not HBM, bounded fc1 (0.65 of the HBM roof); the packed forms
// halve the count.  Same operations, same roundings as the scalar form (1 - th^2 is taken as -(th^2 - 1) with the sign
// folded into negated constants).
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
__device__ __forceinline__ void gelu_fwd2(float2 x, float2& y) {
  const float2 x2 = __fmul2_rn(x, x);
  const float2 u = __fmul2_rn(x, __ffma2_rn(splat2(GELU_B), x2, splat2(GELU_A)));
  const float2 th = make_float2(tanh_approx(u.x), tanh_approx(u.y));
  const float2 hx = __fmul2_rn(splat2(0.5f), x);
  y = __ffma2_rn(hx, th, hx);
}
__device__ __forceinline__ void gelu_fwd_grad2(float2 x, float2& y, float2& dy) {
  const float2 x2 = __fmul2_rn(x, x);
  const float2 u = __fmul2_rn(x, __ffma2_rn(splat2(GELU_B), x2, splat2(GELU_A)));
  const float2 th = make_float2(tanh_approx(u.x), tanh_approx(u.y));
  const float2 hx = __fmul2_rn(splat2(0.5f), x);
  y = __ffma2_rn(hx, th, hx);
  // d/dx [0.5 x (1 + th)] = 0.5 (1 + th) + 0.5 x (1 - th^2) (a + 3 b x^2) = 0.5 (1 + th) + [0.5 x (th^2 - 1)] [-(a + 3 b x^2)]
  const float2 m = __ffma2_rn(th, th, splat2(-1.f));
  const float2 nc = __ffma2_rn(splat2(-3.f * GELU_B), x2, splat2(-GELU_A));
  dy = __ffma2_rn(__fmul2_rn(hx, m), nc, __ffma2_rn(splat2(0.5f), th, splat2(0.5f)));
}

struct Params {
  const float* bias;  // [N] or nullptr (EPI_BIAS / EPI_GELU)
  float* colsum;      // EPI_MUL: [gridDim.x][N] per-CTA partial column sums of out (zeroed by the launcher)
  float* part;        // EPI_F32: [splits][M][N] fp32 partial products
  int M, N, K;
  int a_mn, b_mn;     // operand majors (0 = K-major, 1 = MN-major)
  int has_pre;        // EPI_GELU: also store gelu'(pre-activation) through map_pre
  int splits;         // K splits (EPI_F32 only; 1 otherwise)
  int kb_per_split;   // 64-wide K blocks per split
};

template <int CG, int BN, int EPI>
struct Cfg {
  static constexpr int B_ROWS = BN / CG;                       // rows (K-major) / MN elements (MN-major) of B staged per CTA
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SPT = BN / 64;                          // 64-column slabs per tile
  static constexpr int SPG = (EPI == EPI_GELU || EPI == EPI_MUL) ? 2 : 1;  // staging slabs per quad: out (+ gelu' / multiplier)
  static constexpr int NSLAB = (EPI == EPI_F32) ? 0 : 4 * SPG;
  static constexpr int BAR_BYTES = 1024;
  static constexpr int AVAIL = SMEM_LIMIT - 1024 /*alignment slack*/ - BAR_BYTES - NSLAB * SLAB_BYTES;
  static constexpr int STAGES_RAW = AVAIL / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM = 1024 + STAGES * STAGE_BYTES + NSLAB * SLAB_BYTES + BAR_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;  // double-buffered accumulator (power of two: 256 / 512)
  static constexpr int QUADS_PER_TILE = SPT < 4 ? SPT : 4;     // quads that read one accumulator
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
};

template <int CG, int BN, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                              const __grid_constant__ CUtensorMap map_b,
                                                              const __grid_constant__ CUtensorMap map_out,
                                                              const __grid_constant__ CUtensorMap map_pre,
                                                              const Params p) {
  using C = Cfg<CG, BN, EPI>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // [STAGES][A | B] then the store slabs, every tile 1024-byte aligned in the SHARED address space (swizzle-128B)
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* slabs = tiles + STAGES * C::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(slabs + C::NSLAB * SLAB_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]  (the leader's copy is the one in use)
  uint64_t* mult_full = tmem_empty + 2;   // [4]  EPI_MUL: a quad's multiplier slab has landed
  uint64_t* mult_empty = mult_full + 4;   // [4]  ... and has been consumed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(mult_empty + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / CG, num_clusters = gridDim.x / CG;
  const int m_tiles = (p.M + BM * CG - 1) / (BM * CG), n_tiles = (p.N + BN - 1) / BN;
  const int tiles_mn = m_tiles * n_tiles, num_items = tiles_mn * p.splits;
  const int k_blocks = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; i++) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], CG * 4 * C::QUADS_PER_TILE); }
    for (int i = 0; i < 4; i++) { mbar_init(&mult_full[i], 1); mbar_init(&mult_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 2) {  // one warp (the same warp index in both CTAs of a pair) allocates TMEM and frees it at the end
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "r"(C::TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "r"(C::TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (every CTA stages its own A rows and its share of B) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // completion bytes of BOTH CTAs' loads are counted on the leader's barrier
      uint32_t full_addr[STAGES];
#pragma unroll
      for (int i = 0; i < STAGES; i++) full_addr[i] = (CG == 2) ? mapa_u32(smem_u32(&full[i]), 0) : smem_u32(&full[i]);
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        const int tile = item % tiles_mn, split = item / tiles_mn;
        const int m0 = (tile / n_tiles) * (BM * CG) + (int)rank * BM, n0 = (tile % n_tiles) * BN;
        int n_eff = p.N - n0; n_eff = n_eff >= BN ? BN : ((n_eff + 15) & ~15);
        const int b0 = n0 + (int)rank * (n_eff / CG);  // this CTA's share of the B rows the MMA reads
        const int kb0 = split * p.kb_per_split;
        int kb1 = kb0 + p.kb_per_split; kb1 = kb1 < k_blocks ? kb1 : k_blocks;
        for (int kb = kb0; kb < kb1; kb++) {
          mbar_wait_cluster(&empty[stage], phase ^ 1);
          const uint32_t sa = smem_u32(tiles + stage * C::STAGE_BYTES), sb = sa + A_BYTES;
          if (rank == 0) mbar_expect_tx(&full[stage], CG * C::STAGE_BYTES);
          if (!p.a_mn) {
            tma_load_2d<CG>(sa, &map_a, full_addr[stage], kb * BK, m0);                    // box [64 K][128 rows]
          } else {
            tma_load_2d<CG>(sa, &map_a, full_addr[stage], m0, kb * BK);                    // box [64 M][64 K rows]
            tma_load_2d<CG>(sa + 8192, &map_a, full_addr[stage], m0 + 64, kb * BK);
          }
          if (!p.b_mn) {
            tma_load_2d<CG>(sb, &map_b, full_addr[stage], kb * BK, b0);                    // box [64 K][B_ROWS rows]
          } else {
#pragma unroll
            for (int j = 0; j < C::B_ROWS / 64; j++)
              tma_load_2d<CG>(sb + j * 8192, &map_b, full_addr[stage], b0 + j * 64, kb * BK);  // box [64 N][64 K rows]
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 2 + NUM_EPI_WARPS) {
    // ===================== multiplier stream (EPI_MUL) =====================
    // its own warp: waiting for a staging slot must not hold back the operand prefetch.  A chunk's multiplier goes
    // straight into the slabs the epilogue reads it from.
    if constexpr (EPI == EPI_MUL) {
      if (elect_one()) {
        int uses[4] = {0, 0, 0, 0};  // slabs loaded so far for each quad
        int local = 0;
        for (int item = cluster_id; item < num_items; item += num_clusters, local++) {
          const int tile = item % tiles_mn;
          const int m0 = (tile / n_tiles) * (BM * CG) + (int)rank * BM, n0 = (tile % n_tiles) * BN;
#pragma unroll
          for (int sl = 0; sl < C::SPT; sl++) {
            const int ns = n0 + sl * 64;
            if (ns >= p.N) break;
            const int g = (local * C::SPT + sl) & 3;  // the quad this slab is dealt to
            mbar_wait(&mult_empty[g], ((uses[g] & 1) ^ 1));
            mbar_expect_tx(&mult_full[g], SLAB_BYTES);
            tma_load_2d<1>(smem_u32(slabs + (g * 2 + 1) * SLAB_BYTES), &map_pre, smem_u32(&mult_full[g]), ns, m0);
            uses[g]++;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected thread of the leader CTA) =====================
    if (rank == 0 && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      const uint32_t kstep_a = p.a_mn ? (2048 >> 4) : (32 >> 4), kstep_b = p.b_mn ? (2048 >> 4) : (32 >> 4);
      for (int item = cluster_id; item < num_items; item += num_clusters, local++) {
        const int tile = item % tiles_mn, split = item / tiles_mn;
        const int n0 = (tile % n_tiles) * BN;
        int n_eff = p.N - n0; n_eff = n_eff >= BN ? BN : ((n_eff + 15) & ~15);
        const uint32_t idesc = make_idesc(BM * CG, n_eff, p.a_mn != 0, p.b_mn != 0);
        const int kb0 = split * p.kb_per_split;
        int kb1 = kb0 + p.kb_per_split; kb1 = kb1 < k_blocks ? kb1 : k_blocks;
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        mbar_wait_cluster(&tmem_empty[as], aphase ^ 1);  // both CTAs' epilogues have drained this accumulator stage
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; kb++) {
          mbar_wait_cluster(&full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint32_t sa = smem_u32(tiles + stage * C::STAGE_BYTES);
          const uint64_t adesc = make_smem_desc(sa, p.a_mn != 0), bdesc = make_smem_desc(sa + A_BYTES, p.b_mn != 0);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++)
            umma_bf16<CG>(tmem_d, adesc + k * kstep_a, bdesc + k * kstep_b, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
          umma_commit<CG>(&empty[stage]);  // smem slot reusable (in both CTAs) once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<CG>(&tmem_full[as]);   // accumulator complete -> epilogues of both CTAs
      }
    }
  } else {
    // ===================== epilogue: four independent warp quads; TMEM -> registers -> math -> smem slab -> TMA store ======
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int g = (warp - 2) >> 2;             // quad
    const int r = q * 32 + lane;               // row of the tile owned by this thread
    const bool lead = ((warp - 2) & 3) == 0 && lane == 0;      // first thread of the quad: owns its TMA store stream
    uint8_t* s_out = slabs + g * C::SPG * SLAB_BYTES;          // this quad's staging slab(s)
    uint8_t* s_aux = s_out + SLAB_BYTES;                       // gelu' out / multiplier in (SPG == 2)
    int local = 0, uses = 0;
    for (int item = cluster_id; item < num_items; item += num_clusters, local++) {
      const int tile = item % tiles_mn, split = item / tiles_mn;
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      const int m0 = (tile / n_tiles) * (BM * CG) + (int)rank * BM, n0 = (tile % n_tiles) * BN;
      bool waited = false;
#pragma unroll 1
      for (int sl = 0; sl < C::SPT; sl++) {
        if (((local * C::SPT + sl) & 3) != g) continue;  // slabs are dealt round-robin over the quads
        const int ns = n0 + sl * 64;
        const bool exists = ns < p.N;
        if (!waited) {
          mbar_wait_cluster(&tmem_full[as], aphase);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          waited = true;
        }
        if (exists && EPI != EPI_F32) {
          // this quad's previous TMA store must have finished READING the slab before it is overwritten
          if (lead) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
          quad_bar(g);
          if constexpr (EPI == EPI_MUL) mbar_wait(&mult_full[g], uses & 1);
        }
        if (exists) {
#pragma unroll
          for (int hh = 0; hh < 2; hh++) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + sl * 64 + hh * 32), v);
            const int c0 = ns + hh * 32;
            if constexpr (EPI == EPI_F32) {
              // split-K partial: fp32 straight to the workspace (each thread: 32 consecutive floats of its row)
              const int row = m0 + r;
              if (row < p.M) {
                float* dst = p.part + ((long long)split * p.M + row) * p.N + c0;
#pragma unroll
                for (int j = 0; j < 8; j++)
                  if (c0 + j * 4 < p.N)
                    *reinterpret_cast<float4*>(dst + j * 4) = make_float4(__uint_as_float(v[j * 4]), __uint_as_float(v[j * 4 + 1]),
                                                                           __uint_as_float(v[j * 4 + 2]), __uint_as_float(v[j * 4 + 3]));
              }
            } else {
              uint8_t* so = s_out + r * 128;
              uint8_t* sp = s_aux + r * 128;
#pragma unroll
              for (int j = 0; j < 4; j++) {  // 4 chunks of 8 columns = 16 B
                const int cb = c0 + j * 8;
                const int sw = ((hh * 4 + j) ^ (r & 7)) * 16;  // 128B swizzle: 16-byte chunk index XOR (row mod 8)
                float gq[8];
                if constexpr (EPI == EPI_MUL) {
                  float fm[8];
                  unpack8(*reinterpret_cast<const bf16x8*>(sp + sw), fm);
#pragma unroll
                  for (int t = 0; t < 8; t += 2) {
                    const float2 g2 = __fmul2_rn(make_float2(__uint_as_float(v[j * 8 + t]), __uint_as_float(v[j * 8 + t + 1])),
                                                 make_float2(fm[t], fm[t + 1]));
                    gq[t] = g2.x; gq[t + 1] = g2.y;
                  }
                  *reinterpret_cast<bf16x8*>(so + sw) = pack8(gq);
                } else {
                  float b8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                  if (p.bias && cb < p.N) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + cb));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + cb + 4));
                    b8[0] = b0.x; b8[1] = b0.y; b8[2] = b0.z; b8[3] = b0.w; b8[4] = b1.x; b8[5] = b1.y; b8[6] = b1.z; b8[7] = b1.w;
                  }
                  if constexpr (EPI == EPI_GELU) {
                    if (p.has_pre) {
                      float fq[8];
#pragma unroll
                      for (int t = 0; t < 8; t += 2) {
                        float2 y2, d2;
                        gelu_fwd_grad2(__fadd2_rn(make_float2(__uint_as_float(v[j * 8 + t]), __uint_as_float(v[j * 8 + t + 1])),
                                                  make_float2(b8[t], b8[t + 1])), y2, d2);
                        gq[t] = y2.x; gq[t + 1] = y2.y; fq[t] = d2.x; fq[t + 1] = d2.y;
                      }
                      *reinterpret_cast<bf16x8*>(sp + sw) = pack8(fq);
                    } else {
#pragma unroll
                      for (int t = 0; t < 8; t += 2) {
                        float2 y2;
                        gelu_fwd2(__fadd2_rn(make_float2(__uint_as_float(v[j * 8 + t]), __uint_as_float(v[j * 8 + t + 1])),
                                             make_float2(b8[t], b8[t + 1])), y2);
                        gq[t] = y2.x; gq[t + 1] = y2.y;
                      }
                    }
                  } else {
#pragma unroll
                    for (int t = 0; t < 8; t += 2) {
                      const float2 y2 = __fadd2_rn(make_float2(__uint_as_float(v[j * 8 + t]), __uint_as_float(v[j * 8 + t + 1])),
                                                   make_float2(b8[t], b8[t + 1]));
                      gq[t] = y2.x; gq[t + 1] = y2.y;
                    }
                  }
                  *reinterpret_cast<bf16x8*>(so + sw) = pack8(gq);
                }
              }
            }
          }
        }
        // this warp is done with the accumulator: the MMA warp may start the next-but-one tile once every reader arrived
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(&tmem_empty[as], 0); else mbar_arrive(&tmem_empty[as]); }
        if (exists && EPI != EPI_F32) {
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy smem writes -> visible to TMA
          quad_bar(g);
          if (lead) {
            if constexpr (EPI == EPI_MUL) mbar_arrive(&mult_empty[g]);  // every thread of the quad is past its multiplier reads
            tma_store_2d(&map_out, s_out, ns, m0);
            if constexpr (EPI == EPI_GELU) { if (p.has_pre) tma_store_2d(&map_pre, s_aux, ns, m0); }
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
          }
          if constexpr (EPI == EPI_MUL) {
            // column sums over the staged bf16 slab: thread = (column pair, 32-row group); rows >= M hold zeros.
            // Into this CTA's PRIVATE row: all CTAs adding into one [N] vector serialise in the L2 atomic units.
            const int te = threadIdx.x - 64 - g * 128, cp = te & 31, rg = te >> 5;
            const uint8_t* base = s_out + (cp & 3) * 4;
            const int chunk16 = cp >> 2;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i++) {
              const int row = rg * 32 + i;
              const uint32_t w2 = *reinterpret_cast<const uint32_t*>(base + row * 128 + ((chunk16 ^ (row & 7)) * 16));
              s0 += __uint_as_float(w2 << 16);
              s1 += __uint_as_float(w2 & 0xffff0000u);
            }
            float* dst = p.colsum + (long long)blockIdx.x * p.N;
            const int col = ns + cp * 2;
            if (col < p.N) atomicAdd(dst + col, s0);
            if (col + 1 < p.N) atomicAdd(dst + col + 1, s1);
          }
          uses++;
        }
      }
    }
    if constexpr (EPI != EPI_F32) {
      if (lead) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");  // stores complete before exit
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    if constexpr (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(C::TMEM_COLS));
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(C::TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------------------------------
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
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
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

// out[i] (+)= sum_s part[s][i]  (deterministic fold of the split-K partials)
__global__ void __launch_bounds__(256) split_fold_kernel(const float* __restrict__ part, int splits, long long n4,
                                                         float* __restrict__ out, int accumulate) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  float4 s = accumulate ? reinterpret_cast<const float4*>(out)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < splits; k++) {
    const float4 t = reinterpret_cast<const float4*>(part)[(long long)k * n4 + i];
    s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
  }
  reinterpret_cast<float4*>(out)[i] = s;
}

constexpr int MAX_GRID = 160;  // bound on the persistent grid = rows of the column-sum workspace

struct Call {
  const void* a; const void* b; void* out; const void* aux;  // aux: gelu' out (EPI_GELU) / multiplier in (EPI_MUL)
  Params p;
};

template <int CG, int BN, int EPI>
static int launch_cfg(const Call& c, void* stream, int* grid_out) {
  using C = Cfg<CG, BN, EPI>;
  const Params& p = c.p;
  CUtensorMap ma, mb, mo, mp;
  // A: K-major [M,K] box [64 K][128 rows] | MN-major [K,M] box [64 M][64 K rows]
  if (!(p.a_mn ? make_map(&ma, c.a, p.K, p.M, 64) : make_map(&ma, c.a, p.M, p.K, BM))) return ESVIT_ERR_BAD_ARG;
  // B: K-major [N,K] box [64 K][B_ROWS rows] | MN-major [K,N] box [64 N][64 K rows]
  if (!(p.b_mn ? make_map(&mb, c.b, p.K, p.N, 64) : make_map(&mb, c.b, p.N, p.K, C::B_ROWS))) return ESVIT_ERR_BAD_ARG;
  if (EPI != EPI_F32) {
    if (!make_map(&mo, c.out, p.M, p.N, BM)) return ESVIT_ERR_BAD_ARG;
    if (!make_map(&mp, c.aux ? c.aux : c.out, p.M, p.N, BM)) return ESVIT_ERR_BAD_ARG;
  } else {
    mo = ma; mp = ma;  // unused
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<CG, BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int m_tiles = (p.M + BM * CG - 1) / (BM * CG), n_tiles = (p.N + BN - 1) / BN;
  const long long items = (long long)m_tiles * n_tiles * p.splits;
  int grid = esvit_num_sms();
  if (grid > MAX_GRID) grid = MAX_GRID;
  grid -= grid % CG;
  if ((long long)grid > items * CG) grid = (int)items * CG;
  if (EPI == EPI_MUL) {
    cudaError_t e = cudaMemsetAsync(p.colsum, 0, (size_t)grid * p.N * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
  }
  if (grid_out) *grid_out = grid;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_kernel<CG, BN, EPI>, ma, mb, mo, mp, p);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}

template <int EPI>
static int launch_epi(const Call& c, int cg, int bn, void* stream, int* grid_out) {
  if (cg == 2 && bn == 256) return launch_cfg<2, 256, EPI>(c, stream, grid_out);
  if (cg == 2 && bn == 128) return launch_cfg<2, 128, EPI>(c, stream, grid_out);
  if (cg == 1 && bn == 256) return launch_cfg<1, 256, EPI>(c, stream, grid_out);
  if (cg == 1 && bn == 128) return launch_cfg<1, 128, EPI>(c, stream, grid_out);
  return ESVIT_ERR_BAD_ARG;
}

static int launch(const Call& c, int epi, int cg, int bn, void* stream, int* grid_out = nullptr) {
  switch (epi) {
    case EPI_BIAS: return launch_epi<EPI_BIAS>(c, cg, bn, stream, grid_out);
    case EPI_GELU: return launch_epi<EPI_GELU>(c, cg, bn, stream, grid_out);
    case EPI_MUL: return launch_epi<EPI_MUL>(c, cg, bn, stream, grid_out);
    case EPI_F32: return launch_epi<EPI_F32>(c, cg, bn, stream, grid_out);
  }
  return ESVIT_ERR_BAD_ARG;
}

// Tile shape policy, fitted to scripts/bench_gemm2.py on B200 (profiles/r02_gemm2_microbench.txt); overridable per call
// through `tile` = forced_splits * 10000 + cg * 1000 + bn (0 = automatic).
//   memory-bound shapes (small K, M = all tokens: stages 0 / 1) want MANY small tiles in flight: single CTAs, BN = 128;
//   compute-bound shapes (K >= 512 or enough work per output byte) want CTA pairs with 256 x 256 tiles (B read once per
//   pair, 64 B/clk/SM of operand traffic instead of 128).  `outs` = [M,N] bf16 streams moved besides the operands.
static void pick_tile(long long M, int N, int K, int epi, int tile, int* cg, int* bn) {
  tile %= 10000;
  if (tile > 0) { *cg = tile / 1000; *bn = tile % 1000; return; }
  const int wide = N > 128 ? 256 : 128;
  if (M <= 128) { *cg = 1; *bn = wide; return; }
  const double outs = (epi == EPI_GELU || epi == EPI_MUL) ? 2.0 : 1.0;
  const double inten = (double)N * K / (K + outs * N);          // ~ MACs per byte moved for M >> N, K
  const double lo = outs > 1.5 ? 80.0 : 100.0, hi = outs > 1.5 ? 150.0 : 260.0;
  if (K >= 512 || inten >= hi) { *cg = 2; *bn = wide; }
  else if (inten >= lo) { *cg = 1; *bn = wide; }
  else { *cg = 1; *bn = 128; }
}
// weight gradient (GEMM M = out features, N = in features, contraction over tokens)
static void pick_tile_wgrad(int Nout, int Kin, int tile, int* cg, int* bn) {
  tile %= 10000;
  if (tile > 0) { *cg = tile / 1000; *bn = tile % 1000; return; }
  *bn = Kin > 128 ? 256 : 128;
  *cg = Nout > 128 ? 2 : 1;
}

}  // namespace tg2

// ---- C ABI ----------------------------------------------------------------------------------------------------------
// out[M,N] (bf16) = act( opA(a) . opB(b) + bias[N] )
//   a: a_mn = 0: [M,K] row-major (K contiguous) | a_mn = 1: [K,M] row-major (M contiguous)
//   b: b_mn = 0: [N,K] row-major (nn.Linear weight layout) | b_mn = 1: [K,N] row-major (N contiguous)
//   act 0: identity; act 1: exact GELU (pre != NULL also receives gelu'(pre-activation) for the backward)
//   M, N, K multiples of 8; bias fp32 or NULL.  tile = 0 (automatic) or cta_group * 1000 + BN (1128, 1256, 2128, 2256).
ESVIT_API int esvit_gemm_bf16(const void* a, const void* b, const float* bias, void* out, void* pre, long long M, int N,
                              int K, int a_mn, int b_mn, int act, int tile, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (N % 8) || (K % 8) || M > 0x7fffffffLL || (a_mn && (M % 8)) || act < 0 || act > 1)
    return ESVIT_ERR_BAD_ARG;
  tg2::Call c;
  c.a = a; c.b = b; c.out = out; c.aux = pre;
  c.p.bias = bias; c.p.colsum = nullptr; c.p.part = nullptr; c.p.M = (int)M; c.p.N = N; c.p.K = K;
  c.p.a_mn = a_mn ? 1 : 0; c.p.b_mn = b_mn ? 1 : 0; c.p.has_pre = (act == 1 && pre) ? 1 : 0; c.p.splits = 1;
  c.p.kb_per_split = (K + tg2::BK - 1) / tg2::BK;
  int cg, bn;
  tg2::pick_tile(M, N, K, (act && pre) ? tg2::EPI_GELU : tg2::EPI_BIAS, tile, &cg, &bn);
  return tg2::launch(c, act ? tg2::EPI_GELU : tg2::EPI_BIAS, cg, bn, stream);
}

// out[M,N] (bf16) = (a[M,K] . opB(b)) * mult[M,N];  colsum[N] (fp32) += column sums of out.
// The input-gradient GEMM of fc2 fused with the GELU backward of fc1: a = dy, b = W2 ([K,N] row-major = the Linear weight
// as it lies, b_mn = 1; or a pre-transposed [N,K] copy, b_mn = 0), mult = gelu'(pre-activation) saved by the forward,
// out = d(pre-activation), colsum = fc1 bias gradient (caller zero-fills); ws fp32 [160 * N] scratch.
ESVIT_API int esvit_gemm_mul_colsum2(const void* a, const void* b, const void* mult, void* out, float* colsum, float* ws,
                                     long long M, int N, int K, int b_mn, int tile, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (N % 8) || (K % 8) || M > 0x7fffffffLL || !mult || !colsum || !ws) return ESVIT_ERR_BAD_ARG;
  tg2::Call c;
  c.a = a; c.b = b; c.out = out; c.aux = mult;
  c.p.bias = nullptr; c.p.colsum = ws; c.p.part = nullptr; c.p.M = (int)M; c.p.N = N; c.p.K = K;
  c.p.a_mn = 0; c.p.b_mn = b_mn ? 1 : 0; c.p.has_pre = 0; c.p.splits = 1; c.p.kb_per_split = (K + tg2::BK - 1) / tg2::BK;
  int cg, bn, rows = 0;
  tg2::pick_tile(M, N, K, tg2::EPI_MUL, tile, &cg, &bn);
  const int rc = tg2::launch(c, tg2::EPI_MUL, cg, bn, stream, &rows);
  if (rc != 0) return rc;
  tg2::colsum_fold_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ws, rows, N, colsum);
  ESVIT_LAUNCH_CHECK();
}

// Weight gradient of a Linear: dw[N,K] (fp32) (+)= dy[T,N]^T . x[T,K]   (dy, x bf16 row-major, read as they lie through
// MN-major TMA boxes; any T).  Split over T across the persistent CTAs; partial tiles go to ws (fp32, esvit_gemm_wgrad_ws_floats
// elements) and are folded in a fixed order: bit-reproducible, no atomics.  N, K multiples of 8.
static long long wgrad_ws_floats(int N, int K) {
  const int tiles = ((N + 255) / 256) * ((K + 255) / 256);   // the fewest output tiles any tile shape yields
  long long splits = (2LL * 160 + tiles - 1) / tiles;
  if (splits > 96) splits = 96;                              // bound on the workspace: 96 partial copies of dw
  if (splits < 1) splits = 1;
  return splits * (long long)N * K;
}
// number of fp32 elements of workspace esvit_gemm_wgrad needs for an [N,K] weight (returned as the status-free int)
ESVIT_API int esvit_gemm_wgrad_ws_floats(int N, int K) {
  const long long n = wgrad_ws_floats(N, K);
  return n > 0x7fffffffLL ? -1 : (int)n;
}
ESVIT_API int esvit_gemm_wgrad(const void* dy, const void* x, float* dw, float* ws, long long T, int N, int K,
                               int accumulate, int tile, void* stream) {
  if (T <= 0 || N <= 0 || K <= 0 || (N % 8) || (K % 8) || T > 0x7fffffffLL || !ws) return ESVIT_ERR_BAD_ARG;
  // GEMM view: M = N (out features), N = K (in features), contraction = T
  int cg, bn;
  tg2::pick_tile_wgrad(N, K, tile, &cg, &bn);
  const int m_tiles = (N + tg2::BM * cg - 1) / (tg2::BM * cg), n_tiles = (K + bn - 1) / bn;
  const int tiles = m_tiles * n_tiles;
  const int k_blocks = (int)((T + tg2::BK - 1) / tg2::BK);
  const int clusters = (esvit_num_sms() > tg2::MAX_GRID ? tg2::MAX_GRID : esvit_num_sms()) / cg;
  // one wave of work items: every cluster gets (at most) one K-slice of one output tile, slices of >= 8 K blocks
  int splits = tiles >= clusters ? 1 : clusters / tiles;
  if (splits > k_blocks / 8) splits = k_blocks / 8;
  if (tile >= 10000) splits = tile / 10000;                   // forced (tests / tuning)
  const long long cap = wgrad_ws_floats(N, K) / ((long long)N * K);
  if (splits > cap) splits = (int)cap;
  if (splits > k_blocks) splits = k_blocks;
  if (splits < 1) splits = 1;
  int kbs = (k_blocks + splits - 1) / splits;
  splits = (k_blocks + kbs - 1) / kbs;                         // no empty split
  tg2::Call c;
  c.a = dy; c.b = x; c.out = nullptr; c.aux = nullptr;
  c.p.bias = nullptr; c.p.colsum = nullptr; c.p.part = ws; c.p.M = N; c.p.N = K; c.p.K = (int)T;
  c.p.a_mn = 1; c.p.b_mn = 1; c.p.has_pre = 0; c.p.splits = splits; c.p.kb_per_split = kbs;
  const int rc = tg2::launch(c, tg2::EPI_F32, cg, bn, stream);
  if (rc != 0) return rc;
  const long long n4 = (long long)N * K / 4;
  tg2::split_fold_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ws, splits, n4, dw, accumulate);
  ESVIT_LAUNCH_CHECK();
}

// ---- first-generation entry points (round 1), now served by the same kernel family -------------------------------
// out[M,N] (bf16) = act(a[M,K] @ w[N,K]^T + bias[N]); act 1 = GELU (pre, if not NULL, receives gelu'(pre-activation)).
ESVIT_API int esvit_gemm_bias_act(const void* a, const void* w, const float* bias, void* out, void* pre, long long M,
                                  int N, int K, int act, void* stream) {
  return esvit_gemm_bf16(a, w, bias, out, pre, M, N, K, 0, 0, act, 0, stream);
}
// out[M,N] (bf16) = (a[M,K] @ w[N,K]^T) * mult[M,N];  colsum[N] (fp32) += column sums of out; ws fp32 [160 * N].
ESVIT_API int esvit_gemm_mul_colsum(const void* a, const void* w, const void* mult, void* out, float* colsum, float* ws,
                                    long long M, int N, int K, void* stream) {
  return esvit_gemm_mul_colsum2(a, w, mult, out, colsum, ws, M, N, K, 0, 0, stream);
}
