// Shared helpers for libimpflow_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/impflow_b200.h"

namespace impflow {

// thread-local error string behind impflow_last_error()
void set_error(const char* fmt, ...);
extern long long g_launch_count;
extern int g_pdl;     // programmatic dependent launch of the tile kernels (impflow_set_pdl)

// Device-side gate of the sync-free Broyden loop (conv3_plan.cu).  The host enqueues solver iterations AHEAD of the
// device-side convergence decision; every kernel of such a speculative iteration receives the address of
// state->active and turns into a no-op when the loop has already ended on the device (broyden.py:153 evaluated
// by k_norm_decide).  The launchers of the gated kernels read this thread-local; nullptr = not gated.
extern thread_local const int* g_gate;
struct GateScope {
  const int* prev;
  explicit GateScope(const int* gate) : prev(g_gate) { g_gate = gate; }
  ~GateScope() { g_gate = prev; }
};
__device__ __forceinline__ bool gate_closed(const int* gate) { return gate != nullptr && __ldcg(gate) == 0; }

// Progress record of the sync-free loop, written by the device into MAPPED PINNED host memory (one per iteration,
// index = nstep) and polled by the host: seq == nstep once iteration nstep has been decided.
struct BroydenProgress {
  int seq;
  int active;
};
int broyden_begin_ex(const float* x0, const float* g0, float* xn, float* low_x, float* low_g, float* sample_sq,
                     float* low_sq, float* partial, impflow_broyden_state* state, int B, long long d, int threshold,
                     double eps_scaled, BroydenProgress* progress, void* stream);
// expect_nstep >= 0: k_update only acts if the decision kernel of THIS iteration really ran (state->nstep == expect)
int broyden_step_ex(float* x_old, const float* g_old, const float* xn, const float* gn, float* Ut, float* Vt,
                    float* low_x, float* low_g, float* sample_sq, float* low_sq, float* partial,
                    impflow_broyden_state* state, int B, long long d, int threshold, int gated, int expect_nstep,
                    BroydenProgress* progress, void* stream);

// Programmatic dependent launch (PDL).  A tile kernel launched with the stream-serialisation attribute may be
// scheduled while its predecessor still runs: its prologue (barrier init, TMEM allocation, tensor-map prefetch)
// overlaps the predecessor's tail, and pdl_wait() — executed by every thread before the first global-memory
// access — blocks until the predecessor grid has completed and its writes are visible.  pdl_trigger() in a
// predecessor lets such a dependent be scheduled as soon as all of the predecessor's CTAs have started.  Both are
// no-ops for kernels launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline int pdl_attr(cudaLaunchAttribute* a) {
  if (!g_pdl) return 0;
  a->id = cudaLaunchAttributeProgrammaticStreamSerialization;
  a->val.programmaticStreamSerializationAllowed = 1;
  return 1;
}
// <<<grid, block, 0, s>>> with the PDL attribute; the kernel must call pdl_wait() before its first global access
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = pdl_attr(&attr[0]);
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -1;
  }
  ++g_launch_count;
  return 0;
}

#define IMPFLOW_REQUIRE(cond, ...)        \
  do {                                    \
    if (!(cond)) {                        \
      ::impflow::set_error(__VA_ARGS__);  \
      return -3;                          \
    }                                     \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------
// Activations and their derivatives (reference: lib/layers/base/activations.py:7-12 Sin,
// :64-71 Swish (LipSwish), torch.nn.ReLU).  order = derivative order in x.
// ---------------------------------------------------------------------------------------
constexpr float kTwoPi = 6.283185307179586f;

template <int KIND>
__device__ __forceinline__ float act_eval(float x, int order, float beta) {
  if (KIND == IMPFLOW_ACT_NONE) {
    return order == 0 ? x : (order == 1 ? 1.f : 0.f);
  } else if (KIND == IMPFLOW_ACT_SIN) {
    float s, c;
    sincosf(kTwoPi * x, &s, &c);
    switch (order) {
      case 0: return s / 3.141592653589793f * 0.5f;
      case 1: return c;
      case 2: return -kTwoPi * s;
      default: return -kTwoPi * kTwoPi * c;
    }
  } else if (KIND == IMPFLOW_ACT_RELU) {
    if (order == 0) return x > 0.f ? x : 0.f;
    if (order == 1) return x > 0.f ? 1.f : 0.f;
    return 0.f;
  } else {  // LipSwish: f = x*sigmoid(beta*x)/1.1
    const float bx = beta * x;
    const float s = 1.f / (1.f + expf(-bx));
    const float q = s * (1.f - s);
    const float inv = 1.f / 1.1f;
    switch (order) {
      case 0: return x * s * inv;
      case 1: return (s + bx * q) * inv;
      case 2: return (2.f * beta * q + beta * bx * q * (1.f - 2.f * s)) * inv;
      default: return (3.f * beta * beta * q * (1.f - 2.f * s) +
                       beta * beta * bx * q * (1.f - 6.f * s + 6.f * s * s)) * inv;
    }
  }
}

// LipSwish x*sigmoid(beta*x)/1.1 on the SFU fast paths (ex2.approx + rcp.approx: ~3 ulp, the same order
// as the fp32 accumulation error of the GEMMs around it); used by the GEMM / tile-kernel epilogues.
__device__ __forceinline__ float lipswish_fast(float x, float beta) {
  // e = 2^(-beta x log2 e);  x / (1.1 (1 + e)): e = +inf gives rcp(inf) = 0 and x * 0 = 0 for finite x
  float e, s;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * (-1.4426950408889634f * beta)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(fmaf(e, 1.1f, 1.1f)));
  return x * s;
}

// d/d(beta) of the order-th x-derivative of LipSwish (beta = softplus(raw beta)).
__device__ __forceinline__ float lipswish_dbeta(float x, int order, float beta) {
  const float bx = beta * x;
  const float s = 1.f / (1.f + expf(-bx));
  const float q = s * (1.f - s);
  const float inv = 1.f / 1.1f;
  const float r1 = 1.f - 2.f * s;
  const float r2 = 1.f - 6.f * s + 6.f * s * s;
  switch (order) {
    case 0: return x * x * q * inv;
    case 1: return (2.f * x * q + bx * x * q * r1) * inv;
    default: return (2.f * q + 4.f * bx * q * r1 + bx * bx * q * r2) * inv;
  }
}

__device__ __forceinline__ float act_dispatch(int kind, float x, int order, float beta) {
  switch (kind) {
    case IMPFLOW_ACT_SIN: return act_eval<IMPFLOW_ACT_SIN>(x, order, beta);
    case IMPFLOW_ACT_LIPSWISH: return act_eval<IMPFLOW_ACT_LIPSWISH>(x, order, beta);
    case IMPFLOW_ACT_RELU: return act_eval<IMPFLOW_ACT_RELU>(x, order, beta);
    case IMPFLOW_ACT_MULTIPLIER: return x;      // the stored value IS the derivative factor
    default: return act_eval<IMPFLOW_ACT_NONE>(x, order, beta);
  }
}

// Fused GEMM / col2im epilogue description (see impflow_gemm_nt in the header).
struct Epilogue {
  const float* bias;      // [N] or null
  float* pre_out;         // [M,ldc] or null
  float* act_out;         // [M,ldc] or null
  const float* dmul_pre;  // [M,ldc] or null
  long long ldc;
  int act_kind;
  const float* beta_ptr;  // device scalar softplus(beta) for LipSwish, or null
  float beta;             // filled in by the kernel from beta_ptr
};

__device__ __forceinline__ Epilogue resolve_beta(Epilogue e) {
  e.beta = (e.beta_ptr != nullptr) ? __ldg(e.beta_ptr) : 0.f;
  return e;
}

__device__ __forceinline__ void epilogue_store(const Epilogue& e, long long m, int n, float acc) {
  const long long idx = m * e.ldc + n;
  if (e.dmul_pre != nullptr) {
    if (e.pre_out) e.pre_out[idx] = acc * act_dispatch(e.act_kind, e.dmul_pre[idx], 1, e.beta);
    if (e.act_out) e.act_out[idx] = acc;   // dmul mode: act_out receives the raw accumulator
    return;
  }
  const float v = acc + (e.bias ? e.bias[n] : 0.f);
  if (e.pre_out) e.pre_out[idx] = v;
  if (e.act_out) e.act_out[idx] = act_dispatch(e.act_kind, v, 0, e.beta);
}

}  // namespace impflow
