// Induced 2-norm power iteration for a dense (out,in) matrix, entirely on the device:
// replaces the host loop of mixed_lipschitz.py:85-123 (InducedNormLinear.compute_weight) and
// :276-319 (InducedNormConv2d._compute_weight_1x1) with its per-iteration tolerance test
// (:114-120, including the signed max(u) quirk).  One CTA; u, v live in shared memory.
#include "common.cuh"

namespace impflow {

constexpr int kSnThreads = 512;
constexpr int kSnWarps = kSnThreads / 32;

__device__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < kSnWarps; ++w) t += scratch[w];
  return t;
}
__device__ float block_max(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = scratch[0];
  for (int w = 1; w < kSnWarps; ++w) t = fmaxf(t, scratch[w]);
  return t;
}

// y[r] = sum_c W[r,c] x[c]   (warp per row, coalesced)
__device__ void mat_vec(const float* __restrict__ W, const float* x, float* y, int out_f, int in_f) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < out_f; r += kSnWarps) {
    float acc = 0.f;
    for (int c = lane; c < in_f; c += 32) acc += W[(long long)r * in_f + c] * x[c];
    acc = warp_sum(acc);
    if (lane == 0) y[r] = acc;
  }
  __syncthreads();
}
// y[c] = sum_r W[r,c] x[r]   (thread per column, coalesced across threads)
__device__ void mat_t_vec(const float* __restrict__ W, const float* x, float* y, int out_f, int in_f) {
  for (int c = threadIdx.x; c < in_f; c += kSnThreads) {
    float acc = 0.f;
    for (int r = 0; r < out_f; ++r) acc += W[(long long)r * in_f + c] * x[r];
    y[c] = acc;
  }
  __syncthreads();
}
// F.normalize(p=2, dim=0, eps=1e-12)
__device__ void l2_normalize(float* y, int n, float* scratch) {
  float ss = 0.f;
  for (int i = threadIdx.x; i < n; i += kSnThreads) ss += y[i] * y[i];
  const float nrm = fmaxf(sqrtf(block_sum(ss, scratch)), 1e-12f);
  for (int i = threadIdx.x; i < n; i += kSnThreads) y[i] = y[i] / nrm;
  __syncthreads();
}

__global__ void __launch_bounds__(kSnThreads)
k_sn_power_iter(const float* __restrict__ W, float* __restrict__ u_g, float* __restrict__ v_g,
                float* __restrict__ sigma, int* __restrict__ iters, int out_f, int in_f, int n_iterations,
                float atol, float rtol) {
  extern __shared__ float sm[];
  float* u = sm;
  float* v = u + out_f;
  float* ou = v + in_f;
  float* ov = ou + out_f;
  __shared__ float scratch[kSnWarps];
  for (int i = threadIdx.x; i < out_f; i += kSnThreads) u[i] = u_g[i];
  for (int i = threadIdx.x; i < in_f; i += kSnThreads) v[i] = v_g[i];
  __syncthreads();
  const bool tol_mode = n_iterations < 0;
  const int max_it = tol_mode ? 200 : n_iterations;
  int used = 0;
  for (int it = 0; it < max_it; ++it) {
    for (int i = threadIdx.x; i < out_f; i += kSnThreads) ou[i] = u[i];
    for (int i = threadIdx.x; i < in_f; i += kSnThreads) ov[i] = v[i];
    __syncthreads();
    mat_vec(W, v, u, out_f, in_f);
    l2_normalize(u, out_f, scratch);
    mat_t_vec(W, u, v, out_f, in_f);
    l2_normalize(v, in_f, scratch);
    ++used;
    if (tol_mode) {
      float du = 0.f, dv = 0.f, mu = -INFINITY, mv = -INFINITY;
      for (int i = threadIdx.x; i < out_f; i += kSnThreads) {
        const float t = u[i] - ou[i];
        du += t * t;
        mu = fmaxf(mu, u[i]);
      }
      for (int i = threadIdx.x; i < in_f; i += kSnThreads) {
        const float t = v[i] - ov[i];
        dv += t * t;
        mv = fmaxf(mv, v[i]);
      }
      const float err_u = sqrtf(block_sum(du, scratch)) / sqrtf((float)out_f);
      const float err_v = sqrtf(block_sum(dv, scratch)) / sqrtf((float)in_f);
      const float tol_u = atol + rtol * block_max(mu, scratch);
      const float tol_v = atol + rtol * block_max(mv, scratch);
      if (err_u < tol_u && err_v < tol_v) break;  // uniform: all threads hold the same totals
    }
  }
  // sigma = u^T W v (mixed_lipschitz.py:125)
  mat_vec(W, v, ou, out_f, in_f);
  float s = 0.f;
  for (int i = threadIdx.x; i < out_f; i += kSnThreads) s += u[i] * ou[i];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) {
    sigma[0] = s;
    if (iters) iters[0] = used;
  }
  if (used > 0) {
    for (int i = threadIdx.x; i < out_f; i += kSnThreads) u_g[i] = u[i];
    for (int i = threadIdx.x; i < in_f; i += kSnThreads) v_g[i] = v[i];
  }
}

// W / max(1, sigma/coeff) with sigma read from the device (mixed_lipschitz.py:128-131), and its
// gradient chain  dL/dW = s*G + <G,W> * ds/dsigma * D,  D = dsigma/dW (sigma = <W, D> is linear in W).
__global__ void __launch_bounds__(256)
k_sn_scale(const float* __restrict__ W, const float* __restrict__ sigma, float coeff, float* __restrict__ out,
           float* __restrict__ scale_out, long long n) {
  const float sg = __ldg(sigma);
  const float factor = fmaxf(1.f, sg / coeff);
  if (blockIdx.x == 0 && threadIdx.x == 0 && scale_out != nullptr) scale_out[0] = sg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = W[i] / factor;
}
__global__ void __launch_bounds__(256)
k_sn_scale_grad(const float* __restrict__ G, const float* __restrict__ D, const float* __restrict__ sigma,
                const float* __restrict__ gw_dot, float coeff, float* __restrict__ out, long long n) {
  const float sg = __ldg(sigma);
  const float ratio = sg / coeff;
  const float s = ratio > 1.f ? 1.f / ratio : 1.f;
  const float ds = ratio > 1.f ? -coeff / (sg * sg) : 0.f;
  const float c2 = ds * __ldg(gw_dot);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = s * G[i] + c2 * D[i];
}

}  // namespace impflow

using namespace impflow;

extern "C" int impflow_sn_scale(const float* W, const float* sigma, float coeff, float* out, float* scale_out,
                                long long n, void* stream) {
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  k_sn_scale<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(W, sigma, coeff, out, scale_out, n);
  return check_launch("k_sn_scale");
}

extern "C" int impflow_sn_scale_grad(const float* G, const float* D, const float* sigma, const float* gw_dot,
                                     float coeff, float* out, long long n, void* stream) {
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  k_sn_scale_grad<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(G, D, sigma, gw_dot, coeff, out, n);
  return check_launch("k_sn_scale_grad");
}

extern "C" int impflow_sn_power_iter(const float* W, float* u, float* v, float* sigma, int* iters, int out_f,
                                     int in_f, int n_iterations, float atol, float rtol, void* stream) {
  IMPFLOW_REQUIRE(out_f >= 1 && in_f >= 1, "sn_power_iter: empty matrix");
  const size_t smem = sizeof(float) * 2 * ((size_t)out_f + in_f);
  IMPFLOW_REQUIRE(smem <= 200 * 1024, "sn_power_iter: out+in=%d too large for one CTA", out_f + in_f);
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(k_sn_power_iter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess) {
      set_error("sn_power_iter: cannot set dynamic shared memory");
      return -1;
    }
  }
  k_sn_power_iter<<<1, kSnThreads, smem, (cudaStream_t)stream>>>(W, u, v, sigma, iters, out_f, in_f,
                                                                  n_iterations, atol, rtol);
  return check_launch("k_sn_power_iter");
}
