// ws = 7 (shifted-)window attention core, FORWARD, on tcgen05 tensor cores with TMEM accumulators.
//
// One persistent CTA per SM serves ONE head and walks over PAIRS of windows: the 2 x 64 slots of a pair are the 128 rows
// of one UMMA tile, so every TMEM lane is one query row and the softmax needs no shuffles at all.
//
//   gather warps (4)  cp.async the q / k / v rows of the pair's slots (pad / roll / partition folded into the addressing,
//                     padded slots = the qkv bias, slots >= 49 zero) into 128B-swizzled K-major tiles, 2 stages
//   MMA warp (1 thr)  S = Q K^T     tcgen05.mma M=128 N=128 K=32  -> TMEM   (both windows at once; only the two 64 x 64
//                                                                           diagonal blocks are read back)
//                     O = P V       tcgen05.mma M=128 N=64  K=128 -> TMEM   (P from smem, V read MN-major as it lies)
//   row warps (4)     tcgen05.ld of the thread's own score row (64 columns of its window) -> + rel-pos bias held in
//                     REGISTERS for the whole kernel -> + shift mask -> max / ex2 / sum in registers -> normalised P in bf16
//                     to the swizzled A-operand tile -> later tcgen05.ld of O, bf16, scatter to token order.
// S and O are double-buffered in TMEM (2 x 128 + 2 x 64 columns) and P in shared memory, so the MMAs of pair i+1 run
// under the softmax of pair i; the output epilogue of pair i-1 is deferred behind the softmax of pair i, which hides
// the P V latency.  Same math as window_attn_fwd7_kernel (log2-domain scores, P normalised in fp32 then rounded to
// bf16, natural-log LSE saved for the backward).
#pragma once
#include "wa_common.cuh"

namespace wa {
namespace tc {

constexpr int ROWS = 128;                 // 2 windows x 64 slots
constexpr int TILE_B = ROWS * 128;        // bytes of one operand tile: 128 rows x 128 B (a head's 32 channels use 64 B)
constexpr int STAGE_B = 3 * TILE_B;       // Q | K | V
constexpr int P_B = 2 * TILE_B;           // P: two 64-column K blocks of [128 rows x 128 B]
constexpr int NTHREADS = 32 * 9;          // warps 0-3 rows, 4-7 gather, 8 MMA
constexpr int TMEM_COLS = 512;
constexpr int S_COL = 0, O_COL = 256;     // S[b] at b*128, O[b] at 256 + b*64

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
// bounded spin (~2 s): a mis-programmed pipeline traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return;
    if ((it & 1023u) == 1023u) {
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

struct Smem {
  // byte offsets from the 1024-aligned base
  static constexpr int STAGES = 0;                       // [2][Q | K | V]
  static constexpr int P = 2 * STAGE_B;                  // [2][2 K blocks]
  static constexpr int META = P + 2 * P_B;               // tok [2][128] int, rid [2][128] int
  static constexpr int BARS = META + 2 * 2 * ROWS * 4;   // 14 mbarriers + tmem ptr
  static constexpr int TOTAL = BARS + 256;
};
static size_t fwd7_tc_smem() { return (size_t)Smem::TOTAL + 1024; }

template <bool SHIFT>
__global__ void __launch_bounds__(NTHREADS, 1) window_attn_fwd7_tc_kernel(
    const bf16* __restrict__ qkv, const bf16* __restrict__ qkv_bias, const float* __restrict__ bexp,
    bf16* __restrict__ out, float* __restrict__ lse, Geo g, float scale, int nwin_total) {
  constexpr int WS = 7, NT = 49;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = base + Smem::STAGES;
  uint8_t* pbuf = base + Smem::P;
  int* tokb = reinterpret_cast<int*>(base + Smem::META);  // [2][128]
  int* ridb = tokb + 2 * ROWS;                            // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + Smem::BARS);
  uint64_t* full_qkv = bars;        // [2] count 128 (gather threads)
  uint64_t* empty_qkv = bars + 2;   // [2] count 1   (MMA commit after P V)
  uint64_t* s_full = bars + 4;      // [2] count 1   (MMA commit)
  uint64_t* s_free = bars + 6;      // [2] count 4   (row warps)
  uint64_t* p_full = bars + 8;      // [2] count 4   (row warps)
  uint64_t* o_full = bars + 10;     // [2] count 1   (MMA commit)
  uint64_t* o_free = bars + 12;     // [2] count 4   (row warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);

  const int h = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npairs = (nwin_total + 1) >> 1;
  const int n_items = blockIdx.y < npairs ? (npairs - 1 - (int)blockIdx.y) / (int)gridDim.y + 1 : 0;

  // one-time setup: zero every operand tile (the unused 64 B of each 128 B row feed the N = 64 P V MMA; the off-diagonal
  // P blocks must be 0 and are never written again), barriers, TMEM
  for (int i = threadIdx.x; i < (2 * STAGE_B + 2 * P_B) / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(base)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; i++) {
      mbar_init(&full_qkv[i], 128); mbar_init(&empty_qkv[i], 1); mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&o_free[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // the zero fill is visible to the tensor-core (async) proxy
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= 4 && warp < 8) {
    // ===================== gather warps =====================
    const int t = threadIdx.x - 128;
    const int c16 = t & 3;  // 4 adjacent lanes cover one 64-byte (slot, q|k|v) segment
    uint4 bchunk[3];
#pragma unroll
    for (int part = 0; part < 3; part++)
      bchunk[part] = __ldg(reinterpret_cast<const uint4*>(qkv_bias + part * g.C + h * HD + c16 * 8));
    for (int it = 0; it < n_items; it++) {
      const int b = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int pair = blockIdx.y + it * gridDim.y;
      mbar_wait(&empty_qkv[b], ph ^ 1);
      uint8_t* st = stages + b * STAGE_B;
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        const int r = (t >> 2) + 32 * kk;          // row of the pair tile
        const int w = r >> 6, i = r & 63;
        const int win = 2 * pair + w;
        int tk = -1, rd = 0;
        const bool wexists = win < nwin_total;
        if (wexists && i < NT) {
          const int wx = win % g.nWx, wy = (win / g.nWx) % g.nWy, bb = win / (g.nWx * g.nWy);
          slot_info<WS>(g, bb, wy, wx, i, tk, rd);
        }
        uint8_t* dst = st + r * 128 + ((c16 ^ (r & 7)) * 16);
        if (wexists && i < NT && tk < 0) {         // padded slot: the qkv bias (the reference pads norm1's output with zeros)
#pragma unroll
          for (int part = 0; part < 3; part++) *reinterpret_cast<uint4*>(dst + part * TILE_B) = bchunk[part];
        } else {
          const bf16* src = qkv + (long long)(tk >= 0 ? tk : 0) * 3 * g.C + h * HD + c16 * 8;
          const int nbytes = tk >= 0 ? 16 : 0;     // slots >= 49 / missing second window: zero fill
#pragma unroll
          for (int part = 0; part < 3; part++) cp_async16(dst + part * TILE_B, src + part * g.C, nbytes);
        }
        if (c16 == 0) {
          tokb[b * ROWS + r] = (wexists && i < NT) ? tk : -2;  // -1: padded slot of a real window, -2: no such slot
          ridb[b * ROWS + r] = rd;
        }
      }
      cp_async_commit();
      cp_async_wait<0>();
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> visible to tcgen05.mma
      mbar_arrive(&full_qkv[b]);
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc(128, 128, false);
      const uint32_t idesc_o = make_idesc(128, 64, true);
      for (int it = 0; it <= n_items; it++) {
        if (it < n_items) {  // S(it) = Q K^T
          const int b = it & 1;
          const uint32_t ph = (it >> 1) & 1;
          mbar_wait(&full_qkv[b], ph);
          mbar_wait(&s_free[b], ph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint32_t q = smem_u32(stages + b * STAGE_B);
          const uint64_t adesc = make_desc(q, false), bdesc = make_desc(q + TILE_B, false);
#pragma unroll
          for (int k = 0; k < 2; k++) umma(tmem_base + S_COL + b * 128, adesc + 2 * k, bdesc + 2 * k, idesc_s, k);
          umma_commit(&s_full[b]);
        }
        if (it > 0) {        // O(it-1) = P V
          const int b = (it - 1) & 1;
          const uint32_t ph = ((it - 1) >> 1) & 1;
          mbar_wait(&p_full[b], ph);
          mbar_wait(&o_free[b], ph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint32_t pa = smem_u32(pbuf + b * P_B), va = smem_u32(stages + b * STAGE_B + 2 * TILE_B);
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const uint64_t adesc = make_desc(pa + (k >> 2) * TILE_B, false) + 2 * (k & 3);   // 16 keys = 32 B along K
            const uint64_t bdesc = make_desc(va, true) + (uint64_t)((2048 >> 4) * k);       // 16 key rows of V
            umma(tmem_base + O_COL + b * 64, adesc, bdesc, idesc_o, k);
          }
          umma_commit(&o_full[b]);
          umma_commit(&empty_qkv[b]);  // Q / K / V tiles of this stage are free once these MMAs have read them
        }
      }
    }
  } else {
    // ===================== row warps: thread = query row =====================
    const int r = threadIdx.x;                 // 0..127 = TMEM lane
    const int w = r >> 6, i = r & 63;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float c = scale * LOG2E;
    // rel-pos bias row of this slot (log2 domain, -inf beyond the 49 real keys): registers for the whole kernel
    float bias[64];
    {
      const float4* bp = reinterpret_cast<const float4*>(bexp + (long long)h * 4096 + (i < NT ? i : 0) * 64);
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const float4 v4 = __ldg(bp + j);
        bias[4 * j] = v4.x; bias[4 * j + 1] = v4.y; bias[4 * j + 2] = v4.z; bias[4 * j + 3] = v4.w;
      }
    }
    int tok_prev = -2;
    for (int it = 0; it <= n_items; it++) {
      int tok_cur = -2;
      if (it < n_items) {
        const int b = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int pair = blockIdx.y + it * gridDim.y;
        mbar_wait(&full_qkv[b], ph);           // tok / rid of this stage are written
        tok_cur = tokb[b * ROWS + r];
        int rid_r = 0;
        if (SHIFT) rid_r = ridb[b * ROWS + r];
        mbar_wait(&s_full[b], ph);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        uint32_t v[64];
        tmem_ld32(lane_addr + S_COL + b * 128 + w * 64, v);
        tmem_ld32(lane_addr + S_COL + b * 128 + w * 64 + 32, v + 32);
        tmem_ld_wait();
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[b]);
        float x[64];
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; j++) {
          x[j] = fmaf(__uint_as_float(v[j]), c, bias[j]);
          if (SHIFT) {
            if (j < NT && ridb[b * ROWS + w * 64 + j] != rid_r) x[j] += -100.f * LOG2E;
          }
          m = fmaxf(m, x[j]);
        }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 64; j++) {
          x[j] = ex2(x[j] - m);
          s += x[j];
        }
        const float inv = __fdividef(1.f, s);
        // normalised P (fp32 -> bf16, as the reference under autocast) into the diagonal block of the A-operand tile
        uint8_t* prow = pbuf + b * P_B + w * TILE_B + r * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ch++) {
          float p8[8];
#pragma unroll
          for (int e = 0; e < 8; e++) p8[e] = x[ch * 8 + e] * inv;
          *reinterpret_cast<bf16x8*>(prow + ((ch ^ (r & 7)) * 16)) = pack8(p8);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[b]);
        const int win = 2 * pair + w;
        if (win < nwin_total && i < NT) lse[((long long)win * g.nH + h) * NT + i] = (m + lg2(s)) * LN2;
      }
      if (it > 0) {  // deferred epilogue of the previous pair: O -> bf16 -> token order
        const int b = (it - 1) & 1;
        const uint32_t ph = ((it - 1) >> 1) & 1;
        mbar_wait(&o_full[b], ph);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        uint32_t o[32];
        tmem_ld32(lane_addr + O_COL + b * 64, o);
        tmem_ld_wait();
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[b]);
        if (tok_prev >= 0) {
          bf16* dst = out + (long long)tok_prev * g.C + h * HD;
#pragma unroll
          for (int ch = 0; ch < 4; ch++) {
            float f8[8];
#pragma unroll
            for (int e = 0; e < 8; e++) f8[e] = __uint_as_float(o[ch * 8 + e]);
            *reinterpret_cast<bf16x8*>(dst + ch * 8) = pack8(f8);
          }
        }
      }
      tok_prev = tok_cur;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS));
}

}  // namespace tc
}  // namespace wa
