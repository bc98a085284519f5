// (Shifted-)window attention core, forward and backward, one CTA per (window, head).
//
// Reference: models/swin_transformer.py
//   WindowAttention.forward :120-152      scores = (q*scale) k^T + rel-pos bias (+ shift mask) -> softmax -> @ v
//   SwinTransformerBlock.forward :283-325 zero pad AFTER norm1, roll(-shift), window_partition, ...,
//                                         window_reverse, roll(+shift), crop
//   create_attn_mask :249-272             -100 between tokens of different shift regions
//
// The pad / cyclic roll / window_partition / window_reverse / roll-back / crop copies of the reference are
// folded into the kernel's addressing: a CTA computes, for each of its ws*ws window slots, which token of
// the un-padded [B, H, W] map sits there after pad+roll (or that it is a padded slot, whose q/k/v is the qkv
// bias because the reference pads the *normalised* activations with zeros before the qkv Linear), gathers
// those rows from the token-major qkv tensor and scatters its output rows back to token order.
// The relative-position bias and the -100 shift mask are generated in registers from closed forms
//  - no [nW, N, N] mask or [nH, N, N] bias tensor exists.
//
// Math: bf16 mma.sync m16n8k16 with fp32 accumulate; softmax in fp32 registers; P rounded to bf16 for PV
// (same as the reference under autocast).  head_dim is 32 in every Swin variant.
// The backward recomputes P from the saved log-sum-exp (no [B_, nH, N, N] tensor is saved).
#include "common.cuh"

namespace wa {

constexpr int HD = 32;  // head dim
constexpr int LD = 40;  // smem row stride (bf16 elements) of the q/k/v/dO tiles: 80 B rows -> conflict-free ldmatrix
constexpr int KC = 64;  // key chunk of the backward
constexpr int PLD = KC + 8;  // smem row stride of the P / dS chunk tiles

struct Geo {
  int B, H, W, C, nH, shift, Hp, Wp, nWx, nWy;
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

// gather the q/k/v rows of one (window, head) into smem, adding the qkv bias on the way (the qkv GEMM is bias-free;
// a padded slot holds the bias alone because the reference zero-pads the normalised activations, :287-290).
// qbs: this head's bias, fp32 [3][32].
template <int WS, int NTHREADS>
__device__ __forceinline__ void load_qkv(const Geo& g, const bf16* __restrict__ qkv, const float* qbs, int h,
                                         const int* tok, bf16* Qs, bf16* Ks, bf16* Vs) {
  using C = Cfg<WS>;
  for (int id = threadIdx.x; id < C::KP * 12; id += NTHREADS) {
    const int t = id / 12, rem = id - t * 12, part = rem >> 2, c16 = rem & 3;
    bf16x8 val;
    if (t < C::NT) {
      float f[8];
      const int tk = tok[t];
      if (tk >= 0) {
        unpack8(*reinterpret_cast<const bf16x8*>(qkv + (long long)tk * 3 * g.C + part * g.C + h * HD + c16 * 8), f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; j++) f[j] = 0.f;
      }
      const float* bb = qbs + part * HD + c16 * 8;
#pragma unroll
      for (int j = 0; j < 8; j++) f[j] += bb[j];
      val = pack8(f);
    } else {
      const uint4 z = make_uint4(0, 0, 0, 0);
      val = *reinterpret_cast<const bf16x8*>(&z);
    }
    bf16* dst = (part == 0 ? Qs : (part == 1 ? Ks : Vs)) + t * LD + c16 * 8;
    *reinterpret_cast<bf16x8*>(dst) = val;
  }
}

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

// ------------------------------------------------------------------------------------------------
template <int WS>
__global__ void __launch_bounds__(Cfg<WS>::NW * 32) window_attn_fwd_kernel(
    const bf16* __restrict__ qkv, const float* __restrict__ qkv_bias, const float* __restrict__ bias_table,
    bf16* __restrict__ out, float* __restrict__ lse, Geo g, float scale) {
  using C = Cfg<WS>;
  constexpr int NTHREADS = C::NW * 32;
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + C::KP * LD;
  bf16* Vs = Ks + C::KP * LD;
  float* bt = reinterpret_cast<float*>(Vs + C::KP * LD);
  float* qbs = bt + C::NB;
  int* tok = reinterpret_cast<int*>(qbs + 3 * HD);
  int* rid = tok + C::KP;

  const int win = blockIdx.x, h = blockIdx.y;
  const int wx = win % g.nWx, wy = (win / g.nWx) % g.nWy, b = win / (g.nWx * g.nWy);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (int i = threadIdx.x; i < C::KP; i += NTHREADS) {
    int t = -1, r = 0;
    if (i < C::NT) slot_info<WS>(g, b, wy, wx, i, t, r);
    tok[i] = t;
    rid[i] = r;
  }
  for (int i = threadIdx.x; i < C::NB; i += NTHREADS) bt[i] = bias_table[i * g.nH + h];
  for (int i = threadIdx.x; i < 3 * HD; i += NTHREADS) qbs[i] = qkv_bias[(i / HD) * g.C + h * HD + (i % HD)];
  __syncthreads();
  load_qkv<WS, NTHREADS>(g, qkv, qbs, h, tok, Qs, Ks, Vs);
  __syncthreads();

  for (int mt = warp; mt < C::MT; mt += C::NW) {
    const int r0 = mt * 16;
    uint32_t qa[2][4];
    {
      const bf16* p = Qs + (r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + (lane >> 4) * 8;
      ldsm_x4(qa[0], p);
      ldsm_x4(qa[1], p + 16);
    }
    float acc[C::NT8][4];
#pragma unroll
    for (int nt = 0; nt < C::NT8; nt++) {
      acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      uint32_t kb[4];
      ldsm_x4(kb, Ks + (nt * 8 + (lane & 7)) * LD + (lane >> 3) * 8);
      mma16816(acc[nt], qa[0], kb[0], kb[1]);
      mma16816(acc[nt], qa[1], kb[2], kb[3]);
    }
    const int rA = r0 + (lane >> 2), rB = rA + 8;
    const int ridA = rA < C::NT ? rid[rA] : 0, ridB = rB < C::NT ? rid[rB] : 0;
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < C::NT8; nt++) {
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int row = (e < 2) ? rA : rB;
        const int col = nt * 8 + (lane & 3) * 2 + (e & 1);
        float v;
        if (col < C::NT) {
          if (row < C::NT) {
            v = acc[nt][e] * scale + bt[bias_index<WS>(row, col)];
            if (g.shift > 0 && ((e < 2) ? ridA : ridB) != rid[col]) v += -100.f;
          } else {
            v = 0.f;
          }
        } else {
          v = -INFINITY;
        }
        acc[nt][e] = v;
      }
      m0 = fmaxf(m0, fmaxf(acc[nt][0], acc[nt][1]));
      m1 = fmaxf(m1, fmaxf(acc[nt][2], acc[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < C::NT8; nt++) {
      acc[nt][0] = __expf(acc[nt][0] - m0);
      acc[nt][1] = __expf(acc[nt][1] - m0);
      acc[nt][2] = __expf(acc[nt][2] - m1);
      acc[nt][3] = __expf(acc[nt][3] - m1);
      s0 += acc[nt][0] + acc[nt][1];
      s1 += acc[nt][2] + acc[nt][3];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float i0 = 1.f / s0, i1 = 1.f / s1;
    if ((lane & 3) == 0) {
      float* l = lse + ((long long)win * g.nH + h) * C::NT;
      if (rA < C::NT) l[rA] = m0 + __logf(s0);
      if (rB < C::NT) l[rB] = m1 + __logf(s1);
    }
    float o[4][4];
#pragma unroll
    for (int dt = 0; dt < 4; dt++) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < C::MT; kk++) {
      uint32_t pa[4];
      pa[0] = pack_bf162(acc[2 * kk][0] * i0, acc[2 * kk][1] * i0);
      pa[1] = pack_bf162(acc[2 * kk][2] * i1, acc[2 * kk][3] * i1);
      pa[2] = pack_bf162(acc[2 * kk + 1][0] * i0, acc[2 * kk + 1][1] * i0);
      pa[3] = pack_bf162(acc[2 * kk + 1][2] * i1, acc[2 * kk + 1][3] * i1);
      const bf16* vp = Vs + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + (lane >> 4) * 8;
      uint32_t vb[4];
      ldsm_x4_t(vb, vp);
      mma16816(o[0], pa, vb[0], vb[1]);
      mma16816(o[1], pa, vb[2], vb[3]);
      ldsm_x4_t(vb, vp + 16);
      mma16816(o[2], pa, vb[0], vb[1]);
      mma16816(o[3], pa, vb[2], vb[3]);
    }
    const int tA = rA < C::NT ? tok[rA] : -1, tB = rB < C::NT ? tok[rB] : -1;
#pragma unroll
    for (int dt = 0; dt < 4; dt++) {
      const int d = h * HD + dt * 8 + (lane & 3) * 2;
      if (tA >= 0) *reinterpret_cast<uint32_t*>(out + (long long)tA * g.C + d) = pack_bf162(o[dt][0], o[dt][1]);
      if (tB >= 0) *reinterpret_cast<uint32_t*>(out + (long long)tB * g.C + d) = pack_bf162(o[dt][2], o[dt][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward.  CTA (blockIdx.x = first window, blockIdx.y = head) loops over windows with stride gridDim.x so
// the relative-position-bias and padded-slot (qkv bias) gradients are reduced in shared memory and flushed
// with one global atomic per bin per CTA.
template <int WS>
__global__ void __launch_bounds__(Cfg<WS>::NW * 32) window_attn_bwd_kernel(
    const bf16* __restrict__ qkv, const float* __restrict__ qkv_bias, const float* __restrict__ bias_table,
    const bf16* __restrict__ out, const bf16* __restrict__ dout, const float* __restrict__ lse,
    bf16* __restrict__ dqkv, float* __restrict__ dbias_table, float* __restrict__ dqkv_bias, Geo g, float scale,
    int nwin_total) {
  using C = Cfg<WS>;
  constexpr int NTHREADS = C::NW * 32;
  constexpr int NCH = (C::KP + KC - 1) / KC;  // key chunks
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + C::KP * LD;
  bf16* Vs = Ks + C::KP * LD;
  bf16* dOs = Vs + C::KP * LD;
  bf16* Ps = dOs + C::KP * LD;       // [KP][PLD]  P chunk   (queries x chunk keys)
  bf16* dSs = Ps + C::KP * PLD;      // [KP][PLD]  dS chunk
  float* bt = reinterpret_cast<float*>(dSs + C::KP * PLD);
  float* dbt = bt + C::NB;           // rel-pos bias grad bins (this head)
  float* dqb = dbt + C::NB;          // [3][32] qkv-bias grads of this head (column sums of dq / dk / dv)
  float* qbs = dqb + 3 * HD;         // [3][32] qkv bias of this head
  float* Dsm = qbs + 3 * HD;         // [KP] rowsum(dO * O)
  float* Lsm = Dsm + C::KP;          // [KP] lse
  int* tok = reinterpret_cast<int*>(Lsm + C::KP);
  int* rid = tok + C::KP;

  const int h = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < C::NB; i += NTHREADS) {
    bt[i] = bias_table[i * g.nH + h];
    dbt[i] = 0.f;
  }
  for (int i = threadIdx.x; i < 3 * HD; i += NTHREADS) {
    dqb[i] = 0.f;
    qbs[i] = qkv_bias[(i / HD) * g.C + h * HD + (i % HD)];
  }

  for (int win = blockIdx.x; win < nwin_total; win += gridDim.x) {
    const int wx = win % g.nWx, wy = (win / g.nWx) % g.nWy, b = win / (g.nWx * g.nWy);
    __syncthreads();  // previous window fully consumed (and the init above visible)
    for (int i = threadIdx.x; i < C::KP; i += NTHREADS) {
      int t = -1, r = 0;
      if (i < C::NT) slot_info<WS>(g, b, wy, wx, i, t, r);
      tok[i] = t;
      rid[i] = r;
      Lsm[i] = i < C::NT ? lse[((long long)win * g.nH + h) * C::NT + i] : 0.f;
    }
    __syncthreads();
    load_qkv<WS, NTHREADS>(g, qkv, qbs, h, tok, Qs, Ks, Vs);
    // dO rows (zero for padded slots: their outputs are cropped) and D = rowsum(dO * O)
    for (int id = threadIdx.x; id < C::KP * 4; id += NTHREADS) {
      const int t = id >> 2, c16 = id & 3;
      uint4 dv = make_uint4(0, 0, 0, 0);
      float part = 0.f;
      const int tk = t < C::NT ? tok[t] : -1;
      if (tk >= 0) {
        dv = *reinterpret_cast<const uint4*>(dout + (long long)tk * g.C + h * HD + c16 * 8);
        const uint4 ov = *reinterpret_cast<const uint4*>(out + (long long)tk * g.C + h * HD + c16 * 8);
        float fd[8], fo[8];
        unpack8(*reinterpret_cast<const bf16x8*>(&dv), fd);
        unpack8(*reinterpret_cast<const bf16x8*>(&ov), fo);
#pragma unroll
        for (int j = 0; j < 8; j++) part += fd[j] * fo[j];
      }
      *reinterpret_cast<uint4*>(dOs + t * LD + c16 * 8) = dv;
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      if (c16 == 0) Dsm[t] = part;
    }
    __syncthreads();

    // dQ accumulators of this warp's query tiles (<= 2 tiles per warp)
    constexpr int TPW = (C::MT + C::NW - 1) / C::NW;
    float dq[TPW][4][4];
#pragma unroll
    for (int a = 0; a < TPW; a++)
#pragma unroll
      for (int dt = 0; dt < 4; dt++) dq[a][dt][0] = dq[a][dt][1] = dq[a][dt][2] = dq[a][dt][3] = 0.f;

#pragma unroll 1
    for (int ch = 0; ch < NCH; ch++) {
      const int kc0 = ch * KC;
      // ---- phase 1: per query tile: S, P, dP, dS for this key chunk; dQ += dS K ----
#pragma unroll
      for (int a = 0; a < TPW; a++) {
        const int mt = warp + a * C::NW;
        if (mt < C::MT) {
          const int r0 = mt * 16;
          const int rA = r0 + (lane >> 2), rB = rA + 8;
          uint32_t qa[2][4], da[2][4];
          {
            const int off = (r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + (lane >> 4) * 8;
            ldsm_x4(qa[0], Qs + off);
            ldsm_x4(qa[1], Qs + off + 16);
            ldsm_x4(da[0], dOs + off);
            ldsm_x4(da[1], dOs + off + 16);
          }
          const float lA = Lsm[rA], lB = Lsm[rB], DA = Dsm[rA], DB = Dsm[rB];
          const int ridA = rid[rA], ridB = rid[rB];
          float p[KC / 8][4], ds[KC / 8][4];
#pragma unroll
          for (int nt = 0; nt < KC / 8; nt++) {
            p[nt][0] = p[nt][1] = p[nt][2] = p[nt][3] = 0.f;
            ds[nt][0] = ds[nt][1] = ds[nt][2] = ds[nt][3] = 0.f;
            if (kc0 + nt * 8 < C::KP) {
              uint32_t kb[4];
              ldsm_x4(kb, Ks + (kc0 + nt * 8 + (lane & 7)) * LD + (lane >> 3) * 8);
              mma16816(p[nt], qa[0], kb[0], kb[1]);
              mma16816(p[nt], qa[1], kb[2], kb[3]);
              ldsm_x4(kb, Vs + (kc0 + nt * 8 + (lane & 7)) * LD + (lane >> 3) * 8);
              mma16816(ds[nt], da[0], kb[0], kb[1]);
              mma16816(ds[nt], da[1], kb[2], kb[3]);
#pragma unroll
              for (int e = 0; e < 4; e++) {
                const int row = (e < 2) ? rA : rB;
                const int col = kc0 + nt * 8 + (lane & 3) * 2 + (e & 1);
                float pv = 0.f, dsv = 0.f;
                if (col < C::NT && row < C::NT) {
                  const int bi = bias_index<WS>(row, col);
                  float sv = p[nt][e] * scale + bt[bi];
                  if (g.shift > 0 && ((e < 2) ? ridA : ridB) != rid[col]) sv += -100.f;
                  pv = __expf(sv - ((e < 2) ? lA : lB));
                  dsv = pv * (ds[nt][e] - ((e < 2) ? DA : DB));
                  atomicAdd(&dbt[bi], dsv);
                }
                p[nt][e] = pv;
                ds[nt][e] = dsv;
              }
              // stash bf16 P and dS for the transposed products of phase 2
              const int cc = nt * 8 + (lane & 3) * 2;
              *reinterpret_cast<uint32_t*>(Ps + rA * PLD + cc) = pack_bf162(p[nt][0], p[nt][1]);
              *reinterpret_cast<uint32_t*>(Ps + rB * PLD + cc) = pack_bf162(p[nt][2], p[nt][3]);
              *reinterpret_cast<uint32_t*>(dSs + rA * PLD + cc) = pack_bf162(ds[nt][0], ds[nt][1]);
              *reinterpret_cast<uint32_t*>(dSs + rB * PLD + cc) = pack_bf162(ds[nt][2], ds[nt][3]);
            }
          }
          // dQ += dS_chunk @ K_chunk
#pragma unroll
          for (int kk = 0; kk < KC / 16; kk++) {
            if (kc0 + kk * 16 < C::KP) {
              uint32_t sa[4];
              sa[0] = pack_bf162(ds[2 * kk][0], ds[2 * kk][1]);
              sa[1] = pack_bf162(ds[2 * kk][2], ds[2 * kk][3]);
              sa[2] = pack_bf162(ds[2 * kk + 1][0], ds[2 * kk + 1][1]);
              sa[3] = pack_bf162(ds[2 * kk + 1][2], ds[2 * kk + 1][3]);
              const bf16* kp = Ks + (kc0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + (lane >> 4) * 8;
              uint32_t kb[4];
              ldsm_x4_t(kb, kp);
              mma16816(dq[a][0], sa, kb[0], kb[1]);
              mma16816(dq[a][1], sa, kb[2], kb[3]);
              ldsm_x4_t(kb, kp + 16);
              mma16816(dq[a][2], sa, kb[0], kb[1]);
              mma16816(dq[a][3], sa, kb[2], kb[3]);
            }
          }
        }
      }
      __syncthreads();
      // ---- phase 2: per key tile of the chunk: dV = P^T dO, dK = dS^T Q (over ALL query tiles) ----
      for (int kt = warp; kt < KC / 16; kt += C::NW) {
        const int key0 = kc0 + kt * 16;
        if (key0 >= C::KP) break;
        float dv[4][4], dk[4][4];
#pragma unroll
        for (int dt = 0; dt < 4; dt++) {
          dv[dt][0] = dv[dt][1] = dv[dt][2] = dv[dt][3] = 0.f;
          dk[dt][0] = dk[dt][1] = dk[dt][2] = dk[dt][3] = 0.f;
        }
#pragma unroll 2
        for (int qs = 0; qs < C::MT; qs++) {
          // A(row = key, k = query) from the [query][key] chunk tile, transposed on load
          const int aoff = (qs * 16 + (lane & 7) + (lane >> 4) * 8) * PLD + kt * 16 + ((lane >> 3) & 1) * 8;
          uint32_t pa[4], sa[4];
          ldsm_x4_t(pa, Ps + aoff);
          ldsm_x4_t(sa, dSs + aoff);
          const int boff = (qs * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + (lane >> 4) * 8;
          uint32_t bb[4];
          ldsm_x4_t(bb, dOs + boff);
          mma16816(dv[0], pa, bb[0], bb[1]);
          mma16816(dv[1], pa, bb[2], bb[3]);
          ldsm_x4_t(bb, dOs + boff + 16);
          mma16816(dv[2], pa, bb[0], bb[1]);
          mma16816(dv[3], pa, bb[2], bb[3]);
          ldsm_x4_t(bb, Qs + boff);
          mma16816(dk[0], sa, bb[0], bb[1]);
          mma16816(dk[1], sa, bb[2], bb[3]);
          ldsm_x4_t(bb, Qs + boff + 16);
          mma16816(dk[2], sa, bb[0], bb[1]);
          mma16816(dk[3], sa, bb[2], bb[3]);
        }
        const int kA = key0 + (lane >> 2), kB = kA + 8;
        const int tA = kA < C::NT ? tok[kA] : -1, tB = kB < C::NT ? tok[kB] : -1;
#pragma unroll
        for (int dt = 0; dt < 4; dt++) {
          const int d = h * HD + dt * 8 + (lane & 3) * 2;
          if (tA >= 0) {
            bf16* base = dqkv + (long long)tA * 3 * g.C + d;
            *reinterpret_cast<uint32_t*>(base + g.C) = pack_bf162(dk[dt][0] * scale, dk[dt][1] * scale);
            *reinterpret_cast<uint32_t*>(base + 2 * g.C) = pack_bf162(dv[dt][0], dv[dt][1]);
          }
          if (tB >= 0) {
            bf16* base = dqkv + (long long)tB * 3 * g.C + d;
            *reinterpret_cast<uint32_t*>(base + g.C) = pack_bf162(dk[dt][2] * scale, dk[dt][3] * scale);
            *reinterpret_cast<uint32_t*>(base + 2 * g.C) = pack_bf162(dv[dt][2], dv[dt][3]);
          }
        }
        colsum_to_smem(dk, scale, dqb + HD, lane);
        colsum_to_smem(dv, 1.f, dqb + 2 * HD, lane);
      }
      if (ch + 1 < NCH) __syncthreads();  // P/dS chunk tiles are rewritten by the next chunk
    }
    // ---- dQ rows ----
#pragma unroll
    for (int a = 0; a < TPW; a++) {
      const int mt = warp + a * C::NW;
      if (mt < C::MT) {
        const int rA = mt * 16 + (lane >> 2), rB = rA + 8;
        const int tA = rA < C::NT ? tok[rA] : -1, tB = rB < C::NT ? tok[rB] : -1;
#pragma unroll
        for (int dt = 0; dt < 4; dt++) {
          const int d = h * HD + dt * 8 + (lane & 3) * 2;
          if (tA >= 0)
            *reinterpret_cast<uint32_t*>(dqkv + (long long)tA * 3 * g.C + d) =
                pack_bf162(dq[a][dt][0] * scale, dq[a][dt][1] * scale);
          if (tB >= 0)
            *reinterpret_cast<uint32_t*>(dqkv + (long long)tB * 3 * g.C + d) =
                pack_bf162(dq[a][dt][2] * scale, dq[a][dt][3] * scale);
        }
        colsum_to_smem(dq[a], scale, dqb, lane);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C::NB; i += NTHREADS) atomicAdd(&dbias_table[i * g.nH + h], dbt[i]);
  for (int i = threadIdx.x; i < 3 * HD; i += NTHREADS)
    atomicAdd(&dqkv_bias[(i / HD) * g.C + h * HD + (i % HD)], dqb[i]);
}

// ------------------------------------------------------------------------------------------------
// ws = 7 fast path (KP = 64; one CTA = 4 warps = one (window, head) at a time, PERSISTENT over windows).
//
// Both kernels were instruction-issue bound in their first version (ncu/profiler: ~40 integer/FP instructions per
// score element for the rel-pos index arithmetic, bounds checks and expf), not tensor- or HBM-bound.  Here:
//   * scores live in the log2 domain: s' = acc*(scale*log2e) + bias*log2e (one FMA), P = ex2(s' - m');
//   * the rel-pos bias of this head is expanded ONCE per persistent CTA - forward: straight into the accumulator
//     fragment layout in registers; backward: into a [64][72] fp32 shared-memory table - with -inf in the padded
//     rows/columns, which also replaces every bounds check;
//   * the shift mask is a template flag, so un-shifted blocks carry no mask code.
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int BLD = 72;  // row stride (floats) of the expanded bias table: 72 % 32 == 8 -> conflict-free float2 reads

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

template <bool SHIFT>
__global__ void __launch_bounds__(128, 4) window_attn_fwd7_kernel(
    const bf16* __restrict__ qkv, const float* __restrict__ qkv_bias, const float* __restrict__ bias_table,
    bf16* __restrict__ out, float* __restrict__ lse, Geo g, float scale, int nwin_total) {
  constexpr int WS = 7;
  using C = Cfg<WS>;
  constexpr int NTHREADS = 128;
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + C::KP * LD;
  bf16* Vs = Ks + C::KP * LD;
  float* qbs = reinterpret_cast<float*>(Vs + C::KP * LD);
  int* tok = reinterpret_cast<int*>(qbs + 3 * HD);
  int* rid = tok + C::KP;

  const int h = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = warp * 16, rA = r0 + (lane >> 2), rB = rA + 8;
  for (int i = threadIdx.x; i < 3 * HD; i += NTHREADS) qbs[i] = qkv_bias[(i / HD) * g.C + h * HD + (i % HD)];
  // rel-pos bias of (head h, this warp's 16 query rows) in accumulator-fragment layout, log2 domain
  float breg[C::NT8][4];
#pragma unroll
  for (int nt = 0; nt < C::NT8; nt++)
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int row = (e < 2) ? rA : rB, col = nt * 8 + (lane & 3) * 2 + (e & 1);
      float v = 0.f;
      if (col >= C::NT) v = -INFINITY;
      else if (row < C::NT) v = bias_table[bias_index<WS>(row, col) * g.nH + h] * LOG2E;
      breg[nt][e] = v;
    }
  const float c = scale * LOG2E;
  const int frag_off = (r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + (lane >> 4) * 8;

  for (int win = blockIdx.x; win < nwin_total; win += gridDim.x) {
    const int wx = win % g.nWx, wy = (win / g.nWx) % g.nWy, b = win / (g.nWx * g.nWy);
    __syncthreads();
    if (threadIdx.x < C::KP) {
      int t = -1, r = 0;
      if (threadIdx.x < C::NT) slot_info<WS>(g, b, wy, wx, threadIdx.x, t, r);
      tok[threadIdx.x] = t;
      rid[threadIdx.x] = r;
    }
    __syncthreads();
    load_qkv<WS, NTHREADS>(g, qkv, qbs, h, tok, Qs, Ks, Vs);
    __syncthreads();

    uint32_t qa[2][4];
    ldsm_x4(qa[0], Qs + frag_off);
    ldsm_x4(qa[1], Qs + frag_off + 16);
    float acc[C::NT8][4];
    float m0 = -INFINITY, m1 = -INFINITY;
    int ridA = 0, ridB = 0;
    if (SHIFT) { ridA = rid[rA]; ridB = rid[rB]; }
#pragma unroll
    for (int nt = 0; nt < C::NT8; nt++) {
      acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      uint32_t kb[4];
      ldsm_x4(kb, Ks + (nt * 8 + (lane & 7)) * LD + (lane >> 3) * 8);
      mma16816(acc[nt], qa[0], kb[0], kb[1]);
      mma16816(acc[nt], qa[1], kb[2], kb[3]);
#pragma unroll
      for (int e = 0; e < 4; e++) acc[nt][e] = fmaf(acc[nt][e], c, breg[nt][e]);
      if (SHIFT) {
        const int2 rc = *reinterpret_cast<const int2*>(rid + nt * 8 + (lane & 3) * 2);
        if (ridA != rc.x) acc[nt][0] += -100.f * LOG2E;
        if (ridA != rc.y) acc[nt][1] += -100.f * LOG2E;
        if (ridB != rc.x) acc[nt][2] += -100.f * LOG2E;
        if (ridB != rc.y) acc[nt][3] += -100.f * LOG2E;
      }
      m0 = fmaxf(m0, fmaxf(acc[nt][0], acc[nt][1]));
      m1 = fmaxf(m1, fmaxf(acc[nt][2], acc[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < C::NT8; nt++) {
      acc[nt][0] = ex2(acc[nt][0] - m0);
      acc[nt][1] = ex2(acc[nt][1] - m0);
      acc[nt][2] = ex2(acc[nt][2] - m1);
      acc[nt][3] = ex2(acc[nt][3] - m1);
      s0 += acc[nt][0] + acc[nt][1];
      s1 += acc[nt][2] + acc[nt][3];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float i0 = __frcp_rn(s0), i1 = __frcp_rn(s1);
    if ((lane & 3) == 0) {  // natural-log LSE for the backward
      float* l = lse + ((long long)win * g.nH + h) * C::NT;
      if (rA < C::NT) l[rA] = (m0 + lg2(s0)) * LN2;
      if (rB < C::NT) l[rB] = (m1 + lg2(s1)) * LN2;
    }
    float o[4][4];
#pragma unroll
    for (int dt = 0; dt < 4; dt++) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < C::MT; kk++) {
      uint32_t pa[4];
      pa[0] = pack_bf162(acc[2 * kk][0] * i0, acc[2 * kk][1] * i0);
      pa[1] = pack_bf162(acc[2 * kk][2] * i1, acc[2 * kk][3] * i1);
      pa[2] = pack_bf162(acc[2 * kk + 1][0] * i0, acc[2 * kk + 1][1] * i0);
      pa[3] = pack_bf162(acc[2 * kk + 1][2] * i1, acc[2 * kk + 1][3] * i1);
      const bf16* vp = Vs + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + (lane >> 4) * 8;
      uint32_t vb[4];
      ldsm_x4_t(vb, vp);
      mma16816(o[0], pa, vb[0], vb[1]);
      mma16816(o[1], pa, vb[2], vb[3]);
      ldsm_x4_t(vb, vp + 16);
      mma16816(o[2], pa, vb[0], vb[1]);
      mma16816(o[3], pa, vb[2], vb[3]);
    }
    const int tA = rA < C::NT ? tok[rA] : -1, tB = rB < C::NT ? tok[rB] : -1;
#pragma unroll
    for (int dt = 0; dt < 4; dt++) {
      const int d = h * HD + dt * 8 + (lane & 3) * 2;
      if (tA >= 0) *reinterpret_cast<uint32_t*>(out + (long long)tA * g.C + d) = pack_bf162(o[dt][0], o[dt][1]);
      if (tB >= 0) *reinterpret_cast<uint32_t*>(out + (long long)tB * g.C + d) = pack_bf162(o[dt][2], o[dt][3]);
    }
  }
}

static size_t fwd7_smem() {
  using C = Cfg<7>;
  return (size_t)3 * C::KP * LD * 2 + (size_t)3 * HD * 4 + (size_t)2 * C::KP * 4;
}

// backward: no shared-memory transposition and no atomics in the inner loop.
//   phase A  warp = 16-query tile : S, P, dP, dS  -> dQ = dS K ;  dS also summed into register accumulators
//                                   (this warp's queries x all keys, over all windows) = rel-pos-bias gradient
//   phase B  warp = 16-key tile   : S^T = K Q^T, P^T, dP^T = V dO^T, dS^T recomputed in the transposed layout
//                                   -> dV = P^T dO, dK = dS^T Q straight from the accumulator fragments
// Two __syncthreads per window (smem tile reuse); dqkv-bias gradients are column sums of dQ / dK / dV.
template <bool SHIFT>
__global__ void __launch_bounds__(128, 3) window_attn_bwd7_kernel(
    const bf16* __restrict__ qkv, const float* __restrict__ qkv_bias, const float* __restrict__ bias_table,
    const bf16* __restrict__ out, const bf16* __restrict__ dout, const float* __restrict__ lse,
    bf16* __restrict__ dqkv, float* __restrict__ dbias_table, float* __restrict__ dqkv_bias, Geo g, float scale,
    int nwin_total) {
  constexpr int WS = 7;
  using C = Cfg<WS>;
  constexpr int NTHREADS = 128;
  static_assert(C::KP == 64 && C::NW == 4, "fast path assumes a 64-slot window and 4 warps");
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + C::KP * LD;
  bf16* Vs = Ks + C::KP * LD;
  bf16* dOs = Vs + C::KP * LD;
  float* bm = reinterpret_cast<float*>(dOs + C::KP * LD);  // [64][BLD] expanded bias (log2 domain, -inf padding)
  float* dbt = bm + C::KP * BLD;                           // [NB] bias-gradient bins
  float* dqb = dbt + C::NB + 1;                            // (+1 keeps 8-byte alignment of what follows: NB is odd)
  float* qbs = dqb + 3 * HD;
  float* Dsm = qbs + 3 * HD;
  float* Lsm = Dsm + C::KP;
  int* tok = reinterpret_cast<int*>(Lsm + C::KP);
  int* rid = tok + C::KP;

  const int h = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < C::KP * C::KP; i += NTHREADS) {
    const int row = i >> 6, col = i & 63;
    bm[row * BLD + col] = (row < C::NT && col < C::NT) ? bias_table[bias_index<WS>(row, col) * g.nH + h] * LOG2E
                                                       : -INFINITY;
  }
  for (int i = threadIdx.x; i < C::NB; i += NTHREADS) dbt[i] = 0.f;
  for (int i = threadIdx.x; i < 3 * HD; i += NTHREADS) {
    dqb[i] = 0.f;
    qbs[i] = qkv_bias[(i / HD) * g.C + h * HD + (i % HD)];
  }
  float dsacc[C::NT8][4];
#pragma unroll
  for (int nt = 0; nt < C::NT8; nt++) dsacc[nt][0] = dsacc[nt][1] = dsacc[nt][2] = dsacc[nt][3] = 0.f;

  const float c = scale * LOG2E;
  const int r0 = warp * 16;                       // this warp's query tile (phase A) / key tile (phase B)
  const int rA = r0 + (lane >> 2), rB = rA + 8;
  const int frag_off = (r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + (lane >> 4) * 8;  // A-fragment rows of the tile

  for (int win = blockIdx.x; win < nwin_total; win += gridDim.x) {
    const int wx = win % g.nWx, wy = (win / g.nWx) % g.nWy, b = win / (g.nWx * g.nWy);
    __syncthreads();
    if (threadIdx.x < C::KP) {
      const int i = threadIdx.x;
      int t = -1, r = 0;
      if (i < C::NT) slot_info<WS>(g, b, wy, wx, i, t, r);
      tok[i] = t;
      rid[i] = r;
      Lsm[i] = i < C::NT ? lse[((long long)win * g.nH + h) * C::NT + i] * LOG2E : 0.f;
    }
    __syncthreads();
    load_qkv<WS, NTHREADS>(g, qkv, qbs, h, tok, Qs, Ks, Vs);
    for (int id = threadIdx.x; id < C::KP * 4; id += NTHREADS) {
      const int t = id >> 2, c16 = id & 3;
      uint4 dv = make_uint4(0, 0, 0, 0);
      float part = 0.f;
      const int tk = t < C::NT ? tok[t] : -1;
      if (tk >= 0) {
        dv = *reinterpret_cast<const uint4*>(dout + (long long)tk * g.C + h * HD + c16 * 8);
        const uint4 ov = *reinterpret_cast<const uint4*>(out + (long long)tk * g.C + h * HD + c16 * 8);
        float fd[8], fo[8];
        unpack8(*reinterpret_cast<const bf16x8*>(&dv), fd);
        unpack8(*reinterpret_cast<const bf16x8*>(&ov), fo);
#pragma unroll
        for (int j = 0; j < 8; j++) part += fd[j] * fo[j];
      }
      *reinterpret_cast<uint4*>(dOs + t * LD + c16 * 8) = dv;
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      if (c16 == 0) Dsm[t] = part;
    }
    __syncthreads();

    const int tA = rA < C::NT ? tok[rA] : -1, tB = rB < C::NT ? tok[rB] : -1;
    int ridA = 0, ridB = 0;
    if (SHIFT) { ridA = rid[rA]; ridB = rid[rB]; }
    // ---------------- phase A: rows = queries ----------------
    {
      uint32_t qa[2][4], da[2][4];
      ldsm_x4(qa[0], Qs + frag_off);
      ldsm_x4(qa[1], Qs + frag_off + 16);
      ldsm_x4(da[0], dOs + frag_off);
      ldsm_x4(da[1], dOs + frag_off + 16);
      const float lA = Lsm[rA], lB = Lsm[rB], DA = Dsm[rA], DB = Dsm[rB];
      float dq[4][4];
#pragma unroll
      for (int dt = 0; dt < 4; dt++) dq[dt][0] = dq[dt][1] = dq[dt][2] = dq[dt][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        float ds2[2][4];
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          const int nt = 2 * kk + hf;
          float sacc[4] = {0.f, 0.f, 0.f, 0.f};
          ds2[hf][0] = ds2[hf][1] = ds2[hf][2] = ds2[hf][3] = 0.f;
          uint32_t kb[4];
          const int boff = (nt * 8 + (lane & 7)) * LD + (lane >> 3) * 8;
          ldsm_x4(kb, Ks + boff);
          mma16816(sacc, qa[0], kb[0], kb[1]);
          mma16816(sacc, qa[1], kb[2], kb[3]);
          ldsm_x4(kb, Vs + boff);
          mma16816(ds2[hf], da[0], kb[0], kb[1]);
          mma16816(ds2[hf], da[1], kb[2], kb[3]);
          const int c0 = nt * 8 + (lane & 3) * 2;
          const float2 bA = *reinterpret_cast<const float2*>(bm + rA * BLD + c0);
          const float2 bB = *reinterpret_cast<const float2*>(bm + rB * BLD + c0);
          float sv[4] = {fmaf(sacc[0], c, bA.x) - lA, fmaf(sacc[1], c, bA.y) - lA, fmaf(sacc[2], c, bB.x) - lB,
                         fmaf(sacc[3], c, bB.y) - lB};
          if (SHIFT) {
            const int2 rc = *reinterpret_cast<const int2*>(rid + c0);
            if (ridA != rc.x) sv[0] += -100.f * LOG2E;
            if (ridA != rc.y) sv[1] += -100.f * LOG2E;
            if (ridB != rc.x) sv[2] += -100.f * LOG2E;
            if (ridB != rc.y) sv[3] += -100.f * LOG2E;
          }
          ds2[hf][0] = ex2(sv[0]) * (ds2[hf][0] - DA);
          ds2[hf][1] = ex2(sv[1]) * (ds2[hf][1] - DA);
          ds2[hf][2] = ex2(sv[2]) * (ds2[hf][2] - DB);
          ds2[hf][3] = ex2(sv[3]) * (ds2[hf][3] - DB);
#pragma unroll
          for (int e = 0; e < 4; e++) dsacc[nt][e] += ds2[hf][e];
        }
        uint32_t sa[4];
        sa[0] = pack_bf162(ds2[0][0], ds2[0][1]);
        sa[1] = pack_bf162(ds2[0][2], ds2[0][3]);
        sa[2] = pack_bf162(ds2[1][0], ds2[1][1]);
        sa[3] = pack_bf162(ds2[1][2], ds2[1][3]);
        const bf16* kp = Ks + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + (lane >> 4) * 8;
        uint32_t kb[4];
        ldsm_x4_t(kb, kp);
        mma16816(dq[0], sa, kb[0], kb[1]);
        mma16816(dq[1], sa, kb[2], kb[3]);
        ldsm_x4_t(kb, kp + 16);
        mma16816(dq[2], sa, kb[0], kb[1]);
        mma16816(dq[3], sa, kb[2], kb[3]);
      }
#pragma unroll
      for (int dt = 0; dt < 4; dt++) {
        const int d = h * HD + dt * 8 + (lane & 3) * 2;
        if (tA >= 0)
          *reinterpret_cast<uint32_t*>(dqkv + (long long)tA * 3 * g.C + d) = pack_bf162(dq[dt][0] * scale, dq[dt][1] * scale);
        if (tB >= 0)
          *reinterpret_cast<uint32_t*>(dqkv + (long long)tB * 3 * g.C + d) = pack_bf162(dq[dt][2] * scale, dq[dt][3] * scale);
      }
      colsum_to_smem(dq, scale, dqb, lane);
    }
    // ---------------- phase B: rows = keys (transposed recompute) ----------------
    {
      uint32_t ka[2][4], va[2][4];
      ldsm_x4(ka[0], Ks + frag_off);
      ldsm_x4(ka[1], Ks + frag_off + 16);
      ldsm_x4(va[0], Vs + frag_off);
      ldsm_x4(va[1], Vs + frag_off + 16);
      float dv[4][4], dk[4][4];
#pragma unroll
      for (int dt = 0; dt < 4; dt++) {
        dv[dt][0] = dv[dt][1] = dv[dt][2] = dv[dt][3] = 0.f;
        dk[dt][0] = dk[dt][1] = dk[dt][2] = dk[dt][3] = 0.f;
      }
#pragma unroll
      for (int qq = 0; qq < 4; qq++) {
        float pT[2][4], dsT[2][4];
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          const int nt = 2 * qq + hf;  // 8-query tile
          pT[hf][0] = pT[hf][1] = pT[hf][2] = pT[hf][3] = 0.f;
          dsT[hf][0] = dsT[hf][1] = dsT[hf][2] = dsT[hf][3] = 0.f;
          uint32_t qb[4];
          const int boff = (nt * 8 + (lane & 7)) * LD + (lane >> 3) * 8;
          ldsm_x4(qb, Qs + boff);
          mma16816(pT[hf], ka[0], qb[0], qb[1]);
          mma16816(pT[hf], ka[1], qb[2], qb[3]);
          ldsm_x4(qb, dOs + boff);
          mma16816(dsT[hf], va[0], qb[0], qb[1]);
          mma16816(dsT[hf], va[1], qb[2], qb[3]);
          const int q0 = nt * 8 + (lane & 3) * 2;  // the two query columns of this thread
          const float2 lq = *reinterpret_cast<const float2*>(Lsm + q0);
          const float2 Dq = *reinterpret_cast<const float2*>(Dsm + q0);
          // bias[query][key]: rows q0, q0+1 of the table, columns = this thread's key rows
          float sv[4] = {fmaf(pT[hf][0], c, bm[q0 * BLD + rA]) - lq.x, fmaf(pT[hf][1], c, bm[(q0 + 1) * BLD + rA]) - lq.y,
                         fmaf(pT[hf][2], c, bm[q0 * BLD + rB]) - lq.x, fmaf(pT[hf][3], c, bm[(q0 + 1) * BLD + rB]) - lq.y};
          if (SHIFT) {
            const int2 rq = *reinterpret_cast<const int2*>(rid + q0);
            if (ridA != rq.x) sv[0] += -100.f * LOG2E;
            if (ridA != rq.y) sv[1] += -100.f * LOG2E;
            if (ridB != rq.x) sv[2] += -100.f * LOG2E;
            if (ridB != rq.y) sv[3] += -100.f * LOG2E;
          }
          pT[hf][0] = ex2(sv[0]); pT[hf][1] = ex2(sv[1]); pT[hf][2] = ex2(sv[2]); pT[hf][3] = ex2(sv[3]);
          dsT[hf][0] = pT[hf][0] * (dsT[hf][0] - Dq.x);
          dsT[hf][1] = pT[hf][1] * (dsT[hf][1] - Dq.y);
          dsT[hf][2] = pT[hf][2] * (dsT[hf][2] - Dq.x);
          dsT[hf][3] = pT[hf][3] * (dsT[hf][3] - Dq.y);
        }
        uint32_t pa[4], sa[4];
        pa[0] = pack_bf162(pT[0][0], pT[0][1]);
        pa[1] = pack_bf162(pT[0][2], pT[0][3]);
        pa[2] = pack_bf162(pT[1][0], pT[1][1]);
        pa[3] = pack_bf162(pT[1][2], pT[1][3]);
        sa[0] = pack_bf162(dsT[0][0], dsT[0][1]);
        sa[1] = pack_bf162(dsT[0][2], dsT[0][3]);
        sa[2] = pack_bf162(dsT[1][0], dsT[1][1]);
        sa[3] = pack_bf162(dsT[1][2], dsT[1][3]);
        const int toff = (qq * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + (lane >> 4) * 8;
        uint32_t bb[4];
        ldsm_x4_t(bb, dOs + toff);
        mma16816(dv[0], pa, bb[0], bb[1]);
        mma16816(dv[1], pa, bb[2], bb[3]);
        ldsm_x4_t(bb, dOs + toff + 16);
        mma16816(dv[2], pa, bb[0], bb[1]);
        mma16816(dv[3], pa, bb[2], bb[3]);
        ldsm_x4_t(bb, Qs + toff);
        mma16816(dk[0], sa, bb[0], bb[1]);
        mma16816(dk[1], sa, bb[2], bb[3]);
        ldsm_x4_t(bb, Qs + toff + 16);
        mma16816(dk[2], sa, bb[0], bb[1]);
        mma16816(dk[3], sa, bb[2], bb[3]);
      }
#pragma unroll
      for (int dt = 0; dt < 4; dt++) {
        const int d = h * HD + dt * 8 + (lane & 3) * 2;
        if (tA >= 0) {
          bf16* base = dqkv + (long long)tA * 3 * g.C + d;
          *reinterpret_cast<uint32_t*>(base + g.C) = pack_bf162(dk[dt][0] * scale, dk[dt][1] * scale);
          *reinterpret_cast<uint32_t*>(base + 2 * g.C) = pack_bf162(dv[dt][0], dv[dt][1]);
        }
        if (tB >= 0) {
          bf16* base = dqkv + (long long)tB * 3 * g.C + d;
          *reinterpret_cast<uint32_t*>(base + g.C) = pack_bf162(dk[dt][2] * scale, dk[dt][3] * scale);
          *reinterpret_cast<uint32_t*>(base + 2 * g.C) = pack_bf162(dv[dt][2], dv[dt][3]);
        }
      }
      colsum_to_smem(dk, scale, dqb + HD, lane);
      colsum_to_smem(dv, 1.f, dqb + 2 * HD, lane);
    }
  }
  // flush the register-resident rel-pos-bias gradient of this warp's query rows.  dS was formed with
  // P = ex2(log2-domain score): it is the gradient w.r.t. the natural-domain score, i.e. w.r.t. the table entry.
#pragma unroll
  for (int nt = 0; nt < C::NT8; nt++)
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int row = (e < 2) ? rA : rB;
      const int col = nt * 8 + (lane & 3) * 2 + (e & 1);
      if (row < C::NT && col < C::NT) atomicAdd(&dbt[bias_index<WS>(row, col)], dsacc[nt][e]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < C::NB; i += NTHREADS) atomicAdd(&dbias_table[i * g.nH + h], dbt[i]);
  for (int i = threadIdx.x; i < 3 * HD; i += NTHREADS)
    atomicAdd(&dqkv_bias[(i / HD) * g.C + h * HD + (i % HD)], dqb[i]);
}

static size_t bwd7_smem() {
  using C = Cfg<7>;
  return (size_t)4 * C::KP * LD * 2 + (size_t)(C::KP * BLD + C::NB + 1 + 6 * HD + 2 * C::KP) * 4 + (size_t)2 * C::KP * 4;
}

template <int WS>
size_t fwd_smem() {
  using C = Cfg<WS>;
  return (size_t)3 * C::KP * LD * 2 + (size_t)(C::NB + 3 * HD) * 4 + (size_t)2 * C::KP * 4;
}
template <int WS>
size_t bwd_smem() {
  using C = Cfg<WS>;
  return (size_t)4 * C::KP * LD * 2 + (size_t)2 * C::KP * PLD * 2 + (size_t)(2 * C::NB + 6 * HD + 2 * C::KP) * 4 +
         (size_t)2 * C::KP * 4;
}

static bool make_geo(Geo& g, int B, int H, int W, int C, int nH, int ws, int shift) {
  if (B <= 0 || C != nH * HD || (ws != 7 && ws != 14) || shift < 0 || shift >= ws) return false;
  g.B = B; g.H = H; g.W = W; g.C = C; g.nH = nH; g.shift = shift;
  g.Hp = (H + ws - 1) / ws * ws;
  g.Wp = (W + ws - 1) / ws * ws;
  g.nWy = g.Hp / ws;
  g.nWx = g.Wp / ws;
  return true;
}

}  // namespace wa

// qkv bf16 [B,H,W,3C] = bias-free qkv GEMM output (channel order [q|k|v][head][32]); qkv_bias fp32 [3C] is added
// in-kernel; bias_table fp32 [(2ws-1)^2, nH]; out bf16 [B,H,W,C]; lse fp32 [B*nW, nH, ws*ws]
ESVIT_API int esvit_window_attn_fwd(const void* qkv, const void* qkv_bias, const float* bias_table, void* out,
                                    float* lse, int B, int H, int W, int C, int nH, int ws, int shift, float scale,
                                    void* stream) {
  wa::Geo g;
  if (!wa::make_geo(g, B, H, W, C, nH, ws, shift)) return ESVIT_ERR_BAD_ARG;
  const int nwin = B * g.nWy * g.nWx;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (ws == 7) {
    const size_t smem = wa::fwd7_smem();
    int gx = (esvit_num_sms() * 16 + nH - 1) / nH;  // persistent: ~4 waves of 4 resident CTAs per SM
    if (gx > nwin) gx = nwin;
    if (shift > 0)
      wa::window_attn_fwd7_kernel<true><<<dim3(gx, nH), 128, smem, st>>>(
          (const bf16*)qkv, (const float*)qkv_bias, bias_table, (bf16*)out, lse, g, scale, nwin);
    else
      wa::window_attn_fwd7_kernel<false><<<dim3(gx, nH), 128, smem, st>>>(
          (const bf16*)qkv, (const float*)qkv_bias, bias_table, (bf16*)out, lse, g, scale, nwin);
  } else {
    const size_t smem = wa::fwd_smem<14>();
    e = cudaFuncSetAttribute(wa::window_attn_fwd_kernel<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    wa::window_attn_fwd_kernel<14><<<dim3(nwin, nH), wa::Cfg<14>::NW * 32, smem, st>>>(
        (const bf16*)qkv, (const float*)qkv_bias, bias_table, (bf16*)out, lse, g, scale);
  }
  ESVIT_LAUNCH_CHECK();
}

// dqkv bf16 [B,H,W,3C] (grad of the bias-free GEMM output) is fully written; dbias_table fp32 [(2ws-1)^2, nH] and
// dqkv_bias fp32 [3C] (the COMPLETE qkv-bias gradient: column sums over all window slots, padded ones included)
// are ACCUMULATED into (caller zero-fills).
ESVIT_API int esvit_window_attn_bwd(const void* qkv, const void* qkv_bias, const float* bias_table, const void* out,
                                    const void* dout, const float* lse, void* dqkv, float* dbias_table,
                                    float* dqkv_bias, int B, int H, int W, int C, int nH, int ws, int shift,
                                    float scale, void* stream) {
  wa::Geo g;
  if (!wa::make_geo(g, B, H, W, C, nH, ws, shift)) return ESVIT_ERR_BAD_ARG;
  const int nwin = B * g.nWy * g.nWx;
  cudaStream_t st = (cudaStream_t)stream;
  int gx = (esvit_num_sms() * 8 + nH - 1) / nH;
  if (gx > nwin) gx = nwin;
  if (gx < 1) gx = 1;
  cudaError_t e = cudaSuccess;
  (void)e;
  if (ws == 7) {
    const size_t smem = wa::bwd7_smem();
    gx = (esvit_num_sms() * 12 + nH - 1) / nH;  // 3 CTAs / SM resident, ~4 waves of persistent CTAs
    if (gx > nwin) gx = nwin;
    if (shift > 0)
      wa::window_attn_bwd7_kernel<true><<<dim3(gx, nH), 128, smem, st>>>(
          (const bf16*)qkv, (const float*)qkv_bias, bias_table, (const bf16*)out, (const bf16*)dout, lse,
          (bf16*)dqkv, dbias_table, dqkv_bias, g, scale, nwin);
    else
      wa::window_attn_bwd7_kernel<false><<<dim3(gx, nH), 128, smem, st>>>(
          (const bf16*)qkv, (const float*)qkv_bias, bias_table, (const bf16*)out, (const bf16*)dout, lse,
          (bf16*)dqkv, dbias_table, dqkv_bias, g, scale, nwin);
  } else {
    const size_t smem = wa::bwd_smem<14>();
    e = cudaFuncSetAttribute(wa::window_attn_bwd_kernel<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    wa::window_attn_bwd_kernel<14><<<dim3(gx, nH), wa::Cfg<14>::NW * 32, smem, st>>>(
        (const bf16*)qkv, (const float*)qkv_bias, bias_table, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dqkv,
        dbias_table, dqkv_bias, g, scale, nwin);
  }
  ESVIT_LAUNCH_CHECK();
}
