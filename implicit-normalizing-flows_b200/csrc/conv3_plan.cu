// Native host runtime for the 3x3 / 1x1 / 3x3 conv residual branch of the image flows
// (implicit_flow.py:359-398): one C call per branch evaluation, per power-series chain and per Broyden
// solve, instead of a Python-driven sequence of kernel launches.
//
//   impflow_conv3_forward        nnet(x)                           (implicit_block.py:68-80)
//   impflow_conv3_prepare_vjp    D_l = act'(pre_l), once per saved forward
//   impflow_conv3_vjp            v^T J                             (:199-203, :432-435)
//   impflow_conv3_power_series   w = v + sum_k coeff_k v^T J^k     (Neumann chain, :431-435)
//   impflow_conv3_broyden        whole forward / implicit-backward root solve: branch evaluations, the
//                                residual, the solver algebra and the per-iteration state read-back
//                                (broyden.py:123-193 + implicit_block.py:68-80, 199-207)
//
// Everything works on NHWC "rows" (M = B*H*W rows of c floats): a sample is a contiguous block of
// d = H*W*c floats, which is all the solver's per-sample dot products need.  Shapes with 9c <= 32 and
// C % 256 == 0 take the one-launch tile kernel (branch_fused.cu); the others run im2col-planes +
// three tcgen05 GEMMs + col2im.  The caller owns every buffer (plan->ws included).
#include "common.cuh"

namespace impflow {

static inline size_t pad64(size_t n) { return (n + 63) / 64 * 64; }

struct Conv3Ws {
  float *xin, *x0, *h1, *h2, *Y, *t_rows, *chain_a, *chain_b;
  size_t total;
};

static Conv3Ws carve(float* base, long long M, int c, int C, int k0) {
  Conv3Ws w;
  size_t off = 0;
  auto take = [&](size_t n) {
    float* p = base ? base + off : nullptr;
    off += pad64(n);
    return p;
  };
  w.xin = take((size_t)M * c);
  w.x0 = take(2 * (size_t)M * k0);
  w.h1 = take(2 * (size_t)M * C);
  w.h2 = take(2 * (size_t)M * C);
  w.Y = take((size_t)M * 9 * c);
  w.t_rows = take((size_t)M * c);
  w.chain_a = take((size_t)M * c);
  w.chain_b = take((size_t)M * c);
  w.total = off;
  return w;
}

static inline bool use_tile_kernel(const impflow_conv3_plan* p) {
  return p->allow_fused && p->k0 == 32 && (p->C % 256) == 0 && 9 * p->c <= 32;
}

static int plan_check(const impflow_conv3_plan* p, const char* who) {
  if (p == nullptr || p->ws == nullptr) {
    set_error("%s: plan or workspace missing", who);
    return -3;
  }
  if (p->B < 1 || p->H < 1 || p->W < 1 || p->c < 1 || p->C < 8 || p->k0 < 9 * p->c || (p->k0 % 32) != 0) {
    set_error("%s: bad plan (B=%d H=%d W=%d c=%d C=%d k0=%d)", who, p->B, p->H, p->W, p->c, p->C, p->k0);
    return -3;
  }
  return 0;
}

// y_rows = nnet(x_rows); pre1/pre2 (optional) receive the pre-activations of the two hidden layers.
static int conv3_forward(const impflow_conv3_plan* p, const float* x_rows, float* y_rows, float* pre1, float* pre2,
                         void* stream) {
  const long long M = (long long)p->B * p->H * p->W;
  const Conv3Ws w = carve(p->ws, M, p->c, p->C, p->k0);
  const float* xin = x_rows;
  if (p->act0_kind != IMPFLOW_ACT_NONE) {
    if (impflow_act_mul(x_rows, nullptr, w.xin, M * p->c, p->act0_kind, 0, p->beta0, stream)) return -1;
    xin = w.xin;
  }
  const int N3 = 9 * p->c;
  if (use_tile_kernel(p)) {
    // zero first: the tile kernel then directly follows im2col and its set-up overlaps it (PDL, common.cuh)
    if (p->C > 256 && cudaMemsetAsync(w.Y, 0, sizeof(float) * (size_t)M * N3, (cudaStream_t)stream) != cudaSuccess) {
      set_error("conv3_forward: memset failed");
      return -1;
    }
    if (impflow_im2col3x3(xin, w.x0, p->B, p->H, p->W, p->c, 32, stream)) return -1;
    if (impflow_branch3_tc(w.x0, 32, p->W1f_hi, p->W1f_lo, p->W2f_hi, p->W2f_lo, p->W3f_hi, p->W3f_lo, p->b1, p->b2,
                           nullptr, nullptr, pre1, pre2, w.Y, N3, M, p->C, N3, p->act_kind, p->beta1, p->beta2,
                           stream))
      return -1;
  } else {
    float* x0_hi = w.x0;
    float* x0_lo = w.x0 + (size_t)M * p->k0;
    float* h1_hi = w.h1;
    float* h1_lo = w.h1 + (size_t)M * p->C;
    float* h2_hi = w.h2;
    float* h2_lo = w.h2 + (size_t)M * p->C;
    if (impflow_im2col3x3_split(xin, x0_hi, x0_lo, p->B, p->H, p->W, p->c, p->k0, stream)) return -1;
    if (impflow_gemm_nt_tc(x0_hi, x0_lo, p->k0, p->W1f_hi, p->W1f_lo, p->k0, p->b1, pre1, nullptr, nullptr, h1_hi,
                           h1_lo, p->C, M, p->C, p->k0, p->act_kind, p->beta1, nullptr, stream))
      return -1;
    if (impflow_gemm_nt_tc(h1_hi, h1_lo, p->C, p->W2f_hi, p->W2f_lo, p->C, p->b2, pre2, nullptr, nullptr, h2_hi, h2_lo,
                           p->C, M, p->C, p->C, p->act_kind, p->beta2, nullptr, stream))
      return -1;
    if (impflow_gemm_nt_tc(h2_hi, h2_lo, p->C, p->W3f_hi, p->W3f_lo, p->C, nullptr, w.Y, nullptr, nullptr, nullptr,
                           nullptr, N3, M, N3, p->C, IMPFLOW_ACT_NONE, nullptr, nullptr, stream))
      return -1;
  }
  return impflow_col2im3x3(w.Y, p->B, p->H, p->W, p->c, p->b3, y_rows, nullptr, nullptr, IMPFLOW_ACT_NONE, nullptr,
                           stream);
}

// out_rows = v^T J at the saved point (pre0 = the branch input rows when it has a leading activation).
static int conv3_vjp(const impflow_conv3_plan* p, const float* pre0, const float* d1, const float* d2,
                     const float* v_rows, float* out_rows, void* stream) {
  const long long M = (long long)p->B * p->H * p->W;
  const Conv3Ws w = carve(p->ws, M, p->c, p->C, p->k0);
  const int N3 = 9 * p->c;
  if (use_tile_kernel(p)) {
    if (p->C > 256 && cudaMemsetAsync(w.Y, 0, sizeof(float) * (size_t)M * N3, (cudaStream_t)stream) != cudaSuccess) {
      set_error("conv3_vjp: memset failed");
      return -1;
    }
    if (impflow_im2col3x3(v_rows, w.x0, p->B, p->H, p->W, p->c, 32, stream)) return -1;
    if (impflow_branch3_tc(w.x0, 32, p->W3b_hi, p->W3b_lo, p->W2b_hi, p->W2b_lo, p->W1b_hi, p->W1b_lo, nullptr,
                           nullptr, d2, d1, nullptr, nullptr, w.Y, N3, M, p->C, N3, IMPFLOW_ACT_NONE, nullptr, nullptr,
                           stream))
      return -1;
  } else {
    float* x0_hi = w.x0;
    float* x0_lo = w.x0 + (size_t)M * p->k0;
    float* t3_hi = w.h1;
    float* t3_lo = w.h1 + (size_t)M * p->C;
    float* t2_hi = w.h2;
    float* t2_lo = w.h2 + (size_t)M * p->C;
    if (impflow_im2col3x3_split(v_rows, x0_hi, x0_lo, p->B, p->H, p->W, p->c, p->k0, stream)) return -1;
    if (impflow_gemm_nt_tc(x0_hi, x0_lo, p->k0, p->W3b_hi, p->W3b_lo, p->k0, nullptr, nullptr, nullptr, d2, t3_hi,
                           t3_lo, p->C, M, p->C, p->k0, IMPFLOW_ACT_MULTIPLIER, nullptr, nullptr, stream))
      return -1;
    if (impflow_gemm_nt_tc(t3_hi, t3_lo, p->C, p->W2b_hi, p->W2b_lo, p->C, nullptr, nullptr, nullptr, d1, t2_hi, t2_lo,
                           p->C, M, p->C, p->C, IMPFLOW_ACT_MULTIPLIER, nullptr, nullptr, stream))
      return -1;
    if (impflow_gemm_nt_tc(t2_hi, t2_lo, p->C, p->W1b_hi, p->W1b_lo, p->C, nullptr, w.Y, nullptr, nullptr, nullptr,
                           nullptr, N3, M, N3, p->C, IMPFLOW_ACT_NONE, nullptr, nullptr, stream))
      return -1;
  }
  if (p->act0_kind != IMPFLOW_ACT_NONE)
    return impflow_col2im3x3(w.Y, p->B, p->H, p->W, p->c, nullptr, out_rows, nullptr, pre0, p->act0_kind, p->beta0,
                             stream);
  return impflow_col2im3x3(w.Y, p->B, p->H, p->W, p->c, nullptr, out_rows, nullptr, nullptr, IMPFLOW_ACT_NONE, nullptr,
                           stream);
}

}  // namespace impflow

using namespace impflow;

extern "C" size_t impflow_conv3_workspace_floats(int B, int H, int W, int c, int C, int k0) {
  return carve(nullptr, (long long)B * H * W, c, C, k0).total;
}

extern "C" int impflow_conv3_forward(const impflow_conv3_plan* plan, const float* x_rows, float* y_rows, float* pre1,
                                     float* pre2, void* stream) {
  if (plan_check(plan, "conv3_forward")) return -3;
  return conv3_forward(plan, x_rows, y_rows, pre1, pre2, stream);
}

extern "C" int impflow_conv3_prepare_vjp(const impflow_conv3_plan* plan, const float* pre1, const float* pre2,
                                         float* d1, float* d2, void* stream) {
  if (plan_check(plan, "conv3_prepare_vjp")) return -3;
  const long long n = (long long)plan->B * plan->H * plan->W * plan->C;
  if (impflow_act_mul(pre1, nullptr, d1, n, plan->act_kind, 1, plan->beta1, stream)) return -1;
  return impflow_act_mul(pre2, nullptr, d2, n, plan->act_kind, 1, plan->beta2, stream);
}

extern "C" int impflow_conv3_vjp(const impflow_conv3_plan* plan, const float* pre0, const float* d1, const float* d2,
                                 const float* v_rows, float* out_rows, void* stream) {
  if (plan_check(plan, "conv3_vjp")) return -3;
  IMPFLOW_REQUIRE(plan->act0_kind == IMPFLOW_ACT_NONE || pre0 != nullptr, "conv3_vjp: pre0 missing");
  return conv3_vjp(plan, pre0, d1, d2, v_rows, out_rows, stream);
}

extern "C" int impflow_conv3_power_series(const impflow_conv3_plan* plan, const float* pre0, const float* d1,
                                          const float* d2, const float* v_rows, const double* coeffs, int n,
                                          float* w_rows, void* stream) {
  if (plan_check(plan, "conv3_power_series")) return -3;
  IMPFLOW_REQUIRE(plan->act0_kind == IMPFLOW_ACT_NONE || pre0 != nullptr, "conv3_power_series: pre0 missing");
  const long long M = (long long)plan->B * plan->H * plan->W;
  const long long nel = M * plan->c;
  const Conv3Ws w = carve(plan->ws, M, plan->c, plan->C, plan->k0);
  if (impflow_lincomb3(v_rows, 1.f, nullptr, 0.f, nullptr, 0.f, w_rows, nel, stream)) return -1;
  const float* cur = v_rows;
  float* bufs[2] = {w.chain_a, w.chain_b};
  for (int k = 0; k < n; ++k) {
    float* nxt = bufs[k & 1];
    if (conv3_vjp(plan, pre0, d1, d2, cur, nxt, stream)) return -1;
    if (impflow_lincomb3(w_rows, 1.f, nxt, (float)coeffs[k], nullptr, 0.f, w_rows, nel, stream)) return -1;
    cur = nxt;
  }
  return 0;
}

extern "C" int impflow_conv3_broyden(const impflow_conv3_plan* plan, int mode, const float* rhs_rows,
                                     const float* pre0, const float* d1, const float* d2, float* xa, float* xb,
                                     float* ga, float* gb, float* low_x, float* low_g, float* Ut, float* Vt,
                                     float* sample_sq, float* low_sq, float* partial,
                                     impflow_broyden_state* state_dev, impflow_broyden_state* state_host,
                                     int threshold, double eps_scaled, void* stream) {
  if (plan_check(plan, "conv3_broyden")) return -3;
  IMPFLOW_REQUIRE(mode == 0 || mode == 1, "conv3_broyden: mode must be 0 (forward) or 1 (implicit backward)");
  IMPFLOW_REQUIRE(mode == 0 || plan->act0_kind == IMPFLOW_ACT_NONE || pre0 != nullptr, "conv3_broyden: pre0 missing");
  const long long M = (long long)plan->B * plan->H * plan->W;
  const long long d = (long long)plan->H * plan->W * plan->c;
  const long long nel = M * plan->c;
  const Conv3Ws w = carve(plan->ws, M, plan->c, plan->C, plan->k0);
  cudaStream_t s = (cudaStream_t)stream;
  // residual: forward  g(z) = x_embed - f(z) - z   (implicit_block.py:72)
  //           backward g(v) = v^T J + v - grad      (:199-203)
  auto eval_g = [&](const float* x, float* g) -> int {
    if (mode == 0) {
      if (conv3_forward(plan, x, w.t_rows, nullptr, nullptr, stream)) return -1;
      return impflow_lincomb3(rhs_rows, 1.f, w.t_rows, -1.f, x, -1.f, g, nel, stream);
    }
    if (conv3_vjp(plan, pre0, d1, d2, x, w.t_rows, stream)) return -1;
    return impflow_lincomb3(w.t_rows, 1.f, x, 1.f, rhs_rows, -1.f, g, nel, stream);
  };
  auto read_state = [&]() -> int {
    if (cudaMemcpyAsync(state_host, state_dev, sizeof(impflow_broyden_state), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
      set_error("conv3_broyden: state read-back failed: %s", cudaGetErrorString(cudaGetLastError()));
      return -1;
    }
    return 0;
  };
  float *x_old = xa, *xn = xb, *g_old = ga, *gn = gb;
  if (eval_g(x_old, g_old)) return -1;
  if (impflow_broyden_begin(x_old, g_old, xn, low_x, low_g, sample_sq, low_sq, partial, state_dev, plan->B, d,
                            threshold, eps_scaled, stream))
    return -1;
  if (read_state()) return -1;
  while (state_host->active) {
    if (eval_g(xn, gn)) return -1;
    if (impflow_broyden_step(x_old, g_old, xn, gn, Ut, Vt, low_x, low_g, sample_sq, low_sq, partial, state_dev,
                             plan->B, d, threshold, stream))
      return -1;
    float* t = x_old;      // the kernel wrote the next iterate into the old buffer
    x_old = xn;
    xn = t;
    t = g_old;
    g_old = gn;
    gn = t;
    if (read_state()) return -1;
  }
  return 0;
}
