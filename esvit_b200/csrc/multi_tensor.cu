// Multi-tensor teacher EMA and per-tensor gradient clipping: one launch per <=64 tensors, no host sync.
//
// Reference semantics:
//   EMA   main_esvit.py:587-590   param_k.mul_(m).add_((1 - m) * param_q)   -> fl(fl(k*m) + fl(q*(1-m))), NO fma
//   clip  utils.py:106-115        per-parameter L2 norm; coef = clip / (norm + 1e-6); if coef < 1: grad *= coef
#include "common.cuh"

namespace {

constexpr int MT_MAX = 64;
struct MTList {
  void* a[MT_MAX];
  const void* b[MT_MAX];
  long long n[MT_MAX];
};

__global__ void __launch_bounds__(256) ema_kernel(MTList L, float m, float om) {
  float* k = (float*)L.a[blockIdx.y];
  const float* q = (const float*)L.b[blockIdx.y];
  const long long n = L.n[blockIdx.y];
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
  const bool vec = ((((uintptr_t)k) | ((uintptr_t)q)) & 15) == 0;
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = tid; i < n4; i += nt) {
    float4 a = reinterpret_cast<float4*>(k)[i];
    const float4 b = reinterpret_cast<const float4*>(q)[i];
    a.x = __fadd_rn(__fmul_rn(a.x, m), __fmul_rn(b.x, om));
    a.y = __fadd_rn(__fmul_rn(a.y, m), __fmul_rn(b.y, om));
    a.z = __fadd_rn(__fmul_rn(a.z, m), __fmul_rn(b.z, om));
    a.w = __fadd_rn(__fmul_rn(a.w, m), __fmul_rn(b.w, om));
    reinterpret_cast<float4*>(k)[i] = a;
  }
  for (long long i = n4 * 4 + tid; i < n; i += nt) k[i] = __fadd_rn(__fmul_rn(k[i], m), __fmul_rn(q[i], om));
}

__global__ void __launch_bounds__(256) sumsq_kernel(MTList L, double* __restrict__ sumsq, int base) {
  const float* g = (const float*)L.a[blockIdx.y];
  const long long n = L.n[blockIdx.y];
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
  if ((long long)blockIdx.x * blockDim.x >= n) return;
  const bool vec = (((uintptr_t)g) & 15) == 0;
  const long long n4 = vec ? n / 4 : 0;
  float s = 0.f;
  for (long long i = tid; i < n4; i += nt) {
    const float4 a = reinterpret_cast<const float4*>(g)[i];
    s += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
  }
  for (long long i = n4 * 4 + tid; i < n; i += nt) s += g[i] * g[i];
  __shared__ float sb[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < 8 ? sb[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(&sumsq[base + blockIdx.y], (double)s);
  }
}

__global__ void __launch_bounds__(256) clip_scale_kernel(MTList L, const double* __restrict__ sumsq, int base,
                                                         float clip, float* __restrict__ norms) {
  float* g = (float*)L.a[blockIdx.y];
  const long long n = L.n[blockIdx.y];
  const float norm = (float)sqrt(sumsq[base + blockIdx.y]);
  if (blockIdx.x == 0 && threadIdx.x == 0 && norms) norms[base + blockIdx.y] = norm;
  const float coef = __fdiv_rn(clip, __fadd_rn(norm, 1e-6f));
  if (!(coef < 1.f)) return;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
  const bool vec = (((uintptr_t)g) & 15) == 0;
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = tid; i < n4; i += nt) {
    float4 a = reinterpret_cast<float4*>(g)[i];
    a.x *= coef; a.y *= coef; a.z *= coef; a.w *= coef;
    reinterpret_cast<float4*>(g)[i] = a;
  }
  for (long long i = n4 * 4 + tid; i < n; i += nt) g[i] *= coef;
}

}  // namespace

// teacher[i] = teacher[i] * m + student[i] * (1 - m), fp32, bit-exact with the reference's two ATen ops
ESVIT_API int esvit_ema_multi(void* const* teacher, const void* const* student, const long long* numel, int n,
                              double momentum, void* stream) {
  if (n < 0) return ESVIT_ERR_BAD_ARG;
  const float m = (float)momentum, om = (float)(1.0 - momentum);  // python double -> fp32 scalar, like ATen
  for (int base = 0; base < n; base += MT_MAX) {
    MTList L;
    const int cnt = n - base < MT_MAX ? n - base : MT_MAX;
    for (int i = 0; i < cnt; i++) { L.a[i] = teacher[base + i]; L.b[i] = student[base + i]; L.n[i] = numel[base + i]; }
    ema_kernel<<<dim3(64, cnt), 256, 0, (cudaStream_t)stream>>>(L, m, om);
  }
  ESVIT_LAUNCH_CHECK();
}

// sumsq_ws: double[n] workspace; norms: float[n] (pre-clip norms, may be null)
ESVIT_API int esvit_clip_multi(void* const* grads, const long long* numel, int n, float clip, double* sumsq_ws,
                               float* norms, void* stream) {
  if (n < 0) return ESVIT_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(sumsq_ws, 0, sizeof(double) * (size_t)n, st);
  if (e != cudaSuccess) return (int)e;
  for (int pass = 0; pass < 2; pass++) {
    for (int base = 0; base < n; base += MT_MAX) {
      MTList L;
      const int cnt = n - base < MT_MAX ? n - base : MT_MAX;
      for (int i = 0; i < cnt; i++) { L.a[i] = grads[base + i]; L.b[i] = nullptr; L.n[i] = numel[base + i]; }
      if (pass == 0)
        sumsq_kernel<<<dim3(64, cnt), 256, 0, st>>>(L, sumsq_ws, base);
      else
        clip_scale_kernel<<<dim3(64, cnt), 256, 0, st>>>(L, sumsq_ws, base, clip, norms);
    }
  }
  ESVIT_LAUNCH_CHECK();
}
