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
#include <cstdlib>

#include "wa_common.cuh"
#include "window_attn7.cuh"
#include "window_attn14.cuh"

namespace wa {

// ------------------------------------------------------------------------------------------------
template <int WS>
__global__ void __launch_bounds__(Cfg<WS>::NW * 32) window_attn_fwd_kernel(
    const bf16* __restrict__ qkv, const bf16* __restrict__ qkv_bias, const float* __restrict__ bias_table,
    bf16* __restrict__ out, float* __restrict__ lse, Geo g, float scale) {
  using C = Cfg<WS>;
  constexpr int NTHREADS = C::NW * 32;
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + C::KP * LD;
  bf16* Vs = Ks + C::KP * LD;
  bf16* qbs = Vs + C::KP * LD;  // [3][32] bf16 bias of this head (16-byte aligned: directly behind the tiles)
  float* bt = reinterpret_cast<float*>(qbs + 3 * HD);
  int* tok = reinterpret_cast<int*>(bt + C::NB);
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
    const bf16* __restrict__ qkv, const bf16* __restrict__ qkv_bias, const float* __restrict__ bias_table,
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
  bf16* qbs = dSs + C::KP * PLD;     // [3][32] bf16 bias of this head (16-byte aligned: directly behind the tiles)
  float* bt = reinterpret_cast<float*>(qbs + 3 * HD);
  float* dbt = bt + C::NB;           // rel-pos bias grad bins (this head)
  float* dqb = dbt + C::NB;          // [3][32] qkv-bias grads of this head (column sums of dq / dk / dv)
  float* Dsm = dqb + 3 * HD;         // [KP] rowsum(dO * O)
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
    load_do<WS, NTHREADS>(g, dout, out, h, tok, dOs, Dsm);
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


template <int WS>
size_t fwd_smem() {
  using C = Cfg<WS>;
  return (size_t)3 * C::KP * LD * 2 + (size_t)((C::NB + 3) & ~3) * 4 + (size_t)2 * C::KP * 4 + 3 * HD * 2 + 64;
}
template <int WS>
size_t bwd_smem() {
  using C = Cfg<WS>;
  return (size_t)4 * C::KP * LD * 2 + (size_t)2 * C::KP * PLD * 2 + (size_t)(2 * C::NB + 3 * HD + 2 * C::KP) * 4 +
         (size_t)2 * C::KP * 4 + 3 * HD * 2 + 64;
}

static bool make_geo(Geo& g, int B, int H, int W, int C, int nH, int ws, int shift) {
  if (B <= 0 || C != nH * HD || (ws != 7 && ws != 14) || shift < 0 || shift >= ws) return false;
  g.B = B; g.H = H; g.W = W; g.C = C; g.nH = nH; g.shift = shift;
  g.Hp = (H + ws - 1) / ws * ws;
  g.Wp = (W + ws - 1) / ws * ws;
  g.nWy = g.Hp / ws;
  g.nWx = g.Wp / ws;
  const char* d = getenv("ESVIT_ATTN_DBG");
  g.dbg = d ? atoi(d) : 0;
  return true;
}

// persistent grid: `per_sm` resident CTAs per SM, heads on blockIdx.x.  ESVIT_ATTN_GY (tests) forces a small grid so a
// few windows exercise the multi-window loops.
static int windows_grid(int nwin, int nH, int per_sm) {
  int gy = esvit_num_sms() * per_sm / nH;  // floor: one CTA too many would be a whole extra wave
  if (gy < 1) gy = 1;
  const char* e = getenv("ESVIT_ATTN_GY");
  if (e && atoi(e) > 0) gy = atoi(e);
  return gy > nwin ? nwin : gy;
}
static bool use_generic14() {
  const char* e = getenv("ESVIT_ATTN_GENERIC14");  // A/B switch: the first (generic, ws-templated) kernels
  return e && atoi(e) != 0;
}

template <typename K>
static cudaError_t opt_in_smem(K kernel, size_t smem) {
  return smem > 48 * 1024 ? cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                          : cudaSuccess;
}

}  // namespace wa

// qkv bf16 [B,H,W,3C] = qkv GEMM output INCLUDING its bias (channel order [q|k|v][head][32]); qkv_bias bf16 [3C] is
// what a padded slot holds; bias_table fp32 [(2ws-1)^2, nH]; bias_ws fp32 [nH*4096] workspace (expanded bias, ws=7
// path); out bf16 [B,H,W,C]; lse fp32 [B*nW, nH, ws*ws]
ESVIT_API int esvit_window_attn_fwd(const void* qkv, const void* qkv_bias, const float* bias_table, float* bias_ws,
                                    void* out, float* lse, int B, int H, int W, int C, int nH, int ws, int shift,
                                    float scale, void* stream) {
  wa::Geo g;
  if (!wa::make_geo(g, B, H, W, C, nH, ws, shift)) return ESVIT_ERR_BAD_ARG;
  const int nwin = B * g.nWy * g.nWx;
  cudaStream_t st = (cudaStream_t)stream;
  const bf16* q = (const bf16*)qkv;
  const bf16* qb = (const bf16*)qkv_bias;
  if (ws == 7) {
    if (!bias_ws) return ESVIT_ERR_BAD_ARG;
    wa::expand_bias7_kernel<<<nH, 256, 0, st>>>(bias_table, bias_ws, nH);
    const size_t smem = wa::fwd7_smem();
    const int gx = wa::windows_grid(nwin, nH, 16);  // persistent: ~4 waves of 4 resident CTAs per SM
    if (shift > 0)
      wa::window_attn_fwd7_kernel<true><<<dim3(nH, gx), 128, smem, st>>>(q, qb, bias_ws, (bf16*)out, lse, g, scale, nwin);
    else
      wa::window_attn_fwd7_kernel<false><<<dim3(nH, gx), 128, smem, st>>>(q, qb, bias_ws, (bf16*)out, lse, g, scale, nwin);
  } else if (!wa::use_generic14()) {
    const size_t smem = wa::fwd14_smem();
    cudaError_t e = wa::opt_in_smem(wa::window_attn_fwd14_kernel<true>, smem);
    if (e == cudaSuccess) e = wa::opt_in_smem(wa::window_attn_fwd14_kernel<false>, smem);
    if (e != cudaSuccess) return (int)e;
    const dim3 grid(nH, wa::windows_grid(nwin, nH, 2));
    if (shift > 0)
      wa::window_attn_fwd14_kernel<true><<<grid, wa::T14, smem, st>>>(q, qb, bias_table, (bf16*)out, lse, g, scale, nwin);
    else
      wa::window_attn_fwd14_kernel<false><<<grid, wa::T14, smem, st>>>(q, qb, bias_table, (bf16*)out, lse, g, scale, nwin);
  } else {
    const size_t smem = wa::fwd_smem<14>();
    cudaError_t e = wa::opt_in_smem(wa::window_attn_fwd_kernel<14>, smem);
    if (e != cudaSuccess) return (int)e;
    wa::window_attn_fwd_kernel<14><<<dim3(nwin, nH), wa::Cfg<14>::NW * 32, smem, st>>>(q, qb, bias_table, (bf16*)out, lse,
                                                                                       g, scale);
  }
  ESVIT_LAUNCH_CHECK();
}

// dqkv bf16 [B,H,W,3C] is fully written; dbias_table fp32 [(2ws-1)^2, nH] and dqkv_bias fp32 [3C] (the COMPLETE
// qkv-bias gradient: column sums of dq/dk/dv over all window slots, padded ones included) are ACCUMULATED into
// (caller zero-fills).
ESVIT_API int esvit_window_attn_bwd(const void* qkv, const void* qkv_bias, const float* bias_table, float* bias_ws,
                                    const void* out, const void* dout, const float* lse, void* dqkv,
                                    float* dbias_table, float* dqkv_bias, int B, int H, int W, int C, int nH, int ws,
                                    int shift, float scale, void* stream) {
  wa::Geo g;
  if (!wa::make_geo(g, B, H, W, C, nH, ws, shift)) return ESVIT_ERR_BAD_ARG;
  const int nwin = B * g.nWy * g.nWx;
  cudaStream_t st = (cudaStream_t)stream;
  const bf16* q = (const bf16*)qkv;
  const bf16* qb = (const bf16*)qkv_bias;
  if (ws == 7) {
    if (!bias_ws) return ESVIT_ERR_BAD_ARG;
    wa::expand_bias7_kernel<<<nH, 256, 0, st>>>(bias_table, bias_ws, nH);
    const size_t smem = wa::bwd7_smem();
    cudaError_t e = wa::opt_in_smem(wa::window_attn_bwd7_kernel<true>, smem);
    if (e == cudaSuccess) e = wa::opt_in_smem(wa::window_attn_bwd7_kernel<false>, smem);
    if (e != cudaSuccess) return (int)e;
    const int gx = wa::windows_grid(nwin, nH, 12);  // 3 CTAs / SM resident, ~4 waves of persistent CTAs
    if (shift > 0)
      wa::window_attn_bwd7_kernel<true><<<dim3(nH, gx), 128, smem, st>>>(q, qb, bias_ws, (const bf16*)out,
                                                                         (const bf16*)dout, lse, (bf16*)dqkv,
                                                                         dbias_table, dqkv_bias, g, scale, nwin);
    else
      wa::window_attn_bwd7_kernel<false><<<dim3(nH, gx), 128, smem, st>>>(q, qb, bias_ws, (const bf16*)out,
                                                                          (const bf16*)dout, lse, (bf16*)dqkv,
                                                                          dbias_table, dqkv_bias, g, scale, nwin);
  } else if (!wa::use_generic14()) {
    if (!bias_ws) return ESVIT_ERR_BAD_ARG;
    const size_t smem = wa::bwd14_smem();
    cudaError_t e = wa::opt_in_smem(wa::window_attn_bwd14_kernel<true>, smem);
    if (e == cudaSuccess) e = wa::opt_in_smem(wa::window_attn_bwd14_kernel<false>, smem);
    if (e == cudaSuccess) e = cudaMemsetAsync(bias_ws, 0, (size_t)nH * wa::GACC14 * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    const dim3 grid(nH, wa::windows_grid(nwin, nH, 2));
    if (shift > 0)
      wa::window_attn_bwd14_kernel<true><<<grid, wa::T14, smem, st>>>(q, qb, bias_table, (const bf16*)out, (const bf16*)dout,
                                                                      lse, (bf16*)dqkv, bias_ws, dqkv_bias, g, scale, nwin);
    else
      wa::window_attn_bwd14_kernel<false><<<grid, wa::T14, smem, st>>>(q, qb, bias_table, (const bf16*)out, (const bf16*)dout,
                                                                       lse, (bf16*)dqkv, bias_ws, dqkv_bias, g, scale, nwin);
    wa::fold_dbias14_kernel<<<dim3(27, nH), 192, 0, st>>>(bias_ws, dbias_table, nH);
  } else {
    const size_t smem = wa::bwd_smem<14>();
    cudaError_t e = wa::opt_in_smem(wa::window_attn_bwd_kernel<14>, smem);
    if (e != cudaSuccess) return (int)e;
    int gx = (esvit_num_sms() * 8 + nH - 1) / nH;
    if (gx > nwin) gx = nwin;
    wa::window_attn_bwd_kernel<14><<<dim3(gx, nH), wa::Cfg<14>::NW * 32, smem, st>>>(
        q, qb, bias_table, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dqkv, dbias_table, dqkv_bias, g, scale, nwin);
  }
  ESVIT_LAUNCH_CHECK();
}
