// PatchEmbed: 4x4/stride-4 conv (3 -> E) + LayerNorm as ONE coalesced gather + K=48 FMA + LN kernel.
// Reference: models/swin_transformer.py:537-547 (Conv2d(3,E,4,4) -> flatten(2).transpose(1,2) -> LN).
//
// HBM-bound (reads the fp32 NCHW crops once, writes the fp32 token stream once).  A warp owns 8 horizontally
// adjacent patches: each of the 12 (channel, dy) image rows of that strip is ONE 128-byte coalesced load
// (lane = pixel), the 48 taps of a patch are broadcast from registers with shuffles, lane l produces output
// channels l, l+32, ... from a transposed weight tile in shared memory, and the LN statistics are a warp
// reduction.  The backward recomputes the conv output instead of saving it.
#include "common.cuh"

namespace {

constexpr int PE_K = 48;

template <int EJ>
__device__ __forceinline__ void load_strip(const float* __restrict__ img, int b, int ty, int gx, int H, int W,
                                           int lane, float* px) {
  const int x = gx * 32 + lane;
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int dy = 0; dy < 4; dy++) {
      const int y = ty * 4 + dy;
      px[c * 4 + dy] = (x < W) ? img[(((long long)b * 3 + c) * H + y) * W + x] : 0.f;
    }
}

template <int EJ, int T>
__device__ __forceinline__ void conv_token(const float* px, const float* __restrict__ wsm, const float* bj, int E,
                                           int lane, float* acc) {
#pragma unroll
  for (int j = 0; j < EJ; j++) acc[j] = bj[j];
#pragma unroll
  for (int k = 0; k < PE_K; k++) {
    const float v = __shfl_sync(0xffffffffu, px[(k >> 4) * 4 + ((k >> 2) & 3)], 4 * T + (k & 3));
#pragma unroll
    for (int j = 0; j < EJ; j++) {
      const int c = lane + 32 * j;
      if (c < E) acc[j] += v * wsm[k * E + c];
    }
  }
}

template <int EJ>
__global__ void __launch_bounds__(128) patch_embed_fwd_kernel(
    const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float* __restrict__ out,
    float* __restrict__ mean_o, float* __restrict__ rstd_o, int B, int H, int W, int E) {
  extern __shared__ float wsm[];  // [48][E]
  for (int i = threadIdx.x; i < PE_K * E; i += blockDim.x) {
    const int e = i / PE_K, k = i - e * PE_K;
    wsm[k * E + e] = w[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int Ht = H / 4, Wt = W / 4, G = (Wt + 7) / 8;
  const long long ngroups = (long long)B * Ht * G;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float bj[EJ], gj[EJ], bej[EJ];
#pragma unroll
  for (int j = 0; j < EJ; j++) {
    const int c = lane + 32 * j;
    bj[j] = c < E ? bias[c] : 0.f;
    gj[j] = c < E ? gamma[c] : 0.f;
    bej[j] = c < E ? beta[c] : 0.f;
  }
  const float invE = 1.f / (float)E;
  for (long long g = warp; g < ngroups; g += nwarps) {
    const int gx = (int)(g % G), ty = (int)((g / G) % Ht), b = (int)(g / ((long long)G * Ht));
    float px[12];
    load_strip<EJ>(img, b, ty, gx, H, W, lane, px);
#define PE_TOKEN(T)                                                                      \
  if (gx * 8 + T < Wt) {                                                                 \
    float acc[EJ];                                                                       \
    conv_token<EJ, T>(px, wsm, bj, E, lane, acc);                                        \
    float s = 0.f;                                                                       \
    _Pragma("unroll") for (int j = 0; j < EJ; j++) if (lane + 32 * j < E) s += acc[j];   \
    const float mean = warp_sum(s) * invE;                                               \
    float q = 0.f;                                                                       \
    _Pragma("unroll") for (int j = 0; j < EJ; j++) if (lane + 32 * j < E) {              \
      const float d = acc[j] - mean;                                                     \
      q += d * d;                                                                        \
    }                                                                                    \
    const float rstd = rsqrtf(warp_sum(q) * invE + eps);                                 \
    const long long tok = ((long long)b * Ht + ty) * Wt + gx * 8 + T;                    \
    _Pragma("unroll") for (int j = 0; j < EJ; j++) if (lane + 32 * j < E)                \
        out[tok * E + lane + 32 * j] = (acc[j] - mean) * rstd * gj[j] + bej[j];          \
    if (lane == 0) { mean_o[tok] = mean; rstd_o[tok] = rstd; }                           \
  }
    PE_TOKEN(0) PE_TOKEN(1) PE_TOKEN(2) PE_TOKEN(3) PE_TOKEN(4) PE_TOKEN(5) PE_TOKEN(6) PE_TOKEN(7)
#undef PE_TOKEN
  }
}

// backward: dW[e][k] += dconv[e] * patch[k], dbias += dconv, dgamma += dy*xhat, dbeta += dy
template <int EJ>
__global__ void __launch_bounds__(128, 1) patch_embed_bwd_kernel(
    const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
    const float* __restrict__ gamma, const float* __restrict__ mean_i, const float* __restrict__ rstd_i,
    const float* __restrict__ dout, float* __restrict__ dw, float* __restrict__ dbias, float* __restrict__ dgamma,
    float* __restrict__ dbeta, int B, int H, int W, int E) {
  extern __shared__ float smem[];  // wsm [48][E] | dwsm [48][E] | dvec [3][E]
  float* wsm = smem;
  float* dwsm = smem + PE_K * E;
  float* dvec = dwsm + PE_K * E;
  for (int i = threadIdx.x; i < PE_K * E; i += blockDim.x) {
    const int e = i / PE_K, k = i - e * PE_K;
    wsm[k * E + e] = w[i];
    dwsm[i] = 0.f;
  }
  for (int i = threadIdx.x; i < 3 * E; i += blockDim.x) dvec[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int Ht = H / 4, Wt = W / 4, G = (Wt + 7) / 8;
  const long long ngroups = (long long)B * Ht * G;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float bj[EJ], gj[EJ], adw[EJ][PE_K], adb[EJ], adg[EJ], adbe[EJ];
#pragma unroll
  for (int j = 0; j < EJ; j++) {
    const int c = lane + 32 * j;
    bj[j] = c < E ? bias[c] : 0.f;
    gj[j] = c < E ? gamma[c] : 0.f;
    adb[j] = adg[j] = adbe[j] = 0.f;
#pragma unroll
    for (int k = 0; k < PE_K; k++) adw[j][k] = 0.f;
  }
  const float invE = 1.f / (float)E;
  for (long long g = warp; g < ngroups; g += nwarps) {
    const int gx = (int)(g % G), ty = (int)((g / G) % Ht), b = (int)(g / ((long long)G * Ht));
    float px[12];
    load_strip<EJ>(img, b, ty, gx, H, W, lane, px);
#define PE_TOKEN(T)                                                                               \
  if (gx * 8 + T < Wt) {                                                                          \
    float acc[EJ], gy[EJ], xh[EJ];                                                                \
    conv_token<EJ, T>(px, wsm, bj, E, lane, acc);                                                 \
    const long long tok = ((long long)b * Ht + ty) * Wt + gx * 8 + T;                             \
    const float mean = mean_i[tok], rstd = rstd_i[tok];                                           \
    float s1 = 0.f, s2 = 0.f;                                                                     \
    _Pragma("unroll") for (int j = 0; j < EJ; j++) {                                              \
      const int c = lane + 32 * j;                                                                \
      const float d = c < E ? dout[tok * E + c] : 0.f;                                            \
      xh[j] = c < E ? (acc[j] - mean) * rstd : 0.f;                                               \
      gy[j] = d * gj[j];                                                                          \
      s1 += gy[j];                                                                                \
      s2 += gy[j] * xh[j];                                                                        \
      adg[j] += d * xh[j];                                                                        \
      adbe[j] += d;                                                                               \
    }                                                                                             \
    s1 = warp_sum(s1) * invE;                                                                     \
    s2 = warp_sum(s2) * invE;                                                                     \
    float dc[EJ];                                                                                 \
    _Pragma("unroll") for (int j = 0; j < EJ; j++) {                                              \
      dc[j] = (lane + 32 * j < E) ? rstd * (gy[j] - s1 - xh[j] * s2) : 0.f;                       \
      adb[j] += dc[j];                                                                            \
    }                                                                                             \
    _Pragma("unroll") for (int k = 0; k < PE_K; k++) {                                            \
      const float v = __shfl_sync(0xffffffffu, px[(k >> 4) * 4 + ((k >> 2) & 3)], 4 * T + (k & 3)); \
      _Pragma("unroll") for (int j = 0; j < EJ; j++) adw[j][k] += dc[j] * v;                      \
    }                                                                                             \
  }
    PE_TOKEN(0) PE_TOKEN(1) PE_TOKEN(2) PE_TOKEN(3) PE_TOKEN(4) PE_TOKEN(5) PE_TOKEN(6) PE_TOKEN(7)
#undef PE_TOKEN
  }
#pragma unroll
  for (int j = 0; j < EJ; j++) {
    const int c = lane + 32 * j;
    if (c < E) {
#pragma unroll
      for (int k = 0; k < PE_K; k++) atomicAdd(&dwsm[c * PE_K + k], adw[j][k]);
      atomicAdd(&dvec[c], adb[j]);
      atomicAdd(&dvec[E + c], adg[j]);
      atomicAdd(&dvec[2 * E + c], adbe[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PE_K * E; i += blockDim.x) atomicAdd(&dw[i], dwsm[i]);
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    atomicAdd(&dbias[i], dvec[i]);
    atomicAdd(&dgamma[i], dvec[E + i]);
    atomicAdd(&dbeta[i], dvec[2 * E + i]);
  }
}

}  // namespace

#define PE_DISPATCH(E_, CALL)        \
  if ((E_) <= 32) { CALL(1) }        \
  else if ((E_) <= 64) { CALL(2) }   \
  else if ((E_) <= 96) { CALL(3) }   \
  else if ((E_) <= 128) { CALL(4) }  \
  else return ESVIT_ERR_BAD_ARG;

// img fp32 [B,3,H,W] (H,W multiples of 4); w fp32 [E,3,4,4]; out fp32 [B,(H/4)*(W/4),E]
ESVIT_API int esvit_patch_embed_fwd(const float* img, const float* w, const float* bias, const float* gamma,
                                    const float* beta, float eps, float* out, float* mean, float* rstd, int B, int H,
                                    int W, int E, void* stream) {
  if (H % 4 || W % 4 || B <= 0) return ESVIT_ERR_BAD_ARG;
  const long long ngroups = (long long)B * (H / 4) * ((W / 4 + 7) / 8);
  long long need = (ngroups + 3) / 4, cap = (long long)esvit_num_sms() * 8;
  const int grid = (int)(need < cap ? need : cap);
  const size_t smem = (size_t)PE_K * E * sizeof(float);
#define CALL(EJ) \
  patch_embed_fwd_kernel<EJ><<<grid, 128, smem, (cudaStream_t)stream>>>(img, w, bias, gamma, beta, eps, out, mean, rstd, B, H, W, E);
  PE_DISPATCH(E, CALL)
#undef CALL
  ESVIT_LAUNCH_CHECK();
}

// dw [E,48], dbias/dgamma/dbeta [E] are ACCUMULATED into (caller zero-fills)
ESVIT_API int esvit_patch_embed_bwd(const float* img, const float* w, const float* bias, const float* gamma,
                                    const float* mean, const float* rstd, const float* dout, float* dw, float* dbias,
                                    float* dgamma, float* dbeta, int B, int H, int W, int E, void* stream) {
  if (H % 4 || W % 4 || B <= 0) return ESVIT_ERR_BAD_ARG;
  const long long ngroups = (long long)B * (H / 4) * ((W / 4 + 7) / 8);
  long long need = (ngroups + 3) / 4, cap = (long long)esvit_num_sms() * 2;
  const int grid = (int)(need < cap ? need : cap);
  const size_t smem = (size_t)(2 * PE_K + 3) * E * sizeof(float);
#define CALL(EJ)                                                                                                   \
  {                                                                                                                \
    cudaError_t e = cudaFuncSetAttribute(patch_embed_bwd_kernel<EJ>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                         (int)smem);                                                               \
    if (e != cudaSuccess) return (int)e;                                                                           \
    patch_embed_bwd_kernel<EJ><<<grid, 128, smem, (cudaStream_t)stream>>>(img, w, bias, gamma, mean, rstd, dout,   \
                                                                          dw, dbias, dgamma, dbeta, B, H, W, E);   \
  }
  PE_DISPATCH(E, CALL)
#undef CALL
  ESVIT_LAUNCH_CHECK();
}
