// Induced 2-norm power iteration of a 3x3 / stride 1 / pad 1 convolution acting on ONE h x w image,
// entirely on the device in ONE cooperative launch: replaces the host loop of
// mixed_lipschitz.py:328-386 (InducedNormConv2d._compute_weight_kxk: per iteration a 1-sample conv2d,
// a conv_transpose2d, two normalisations, four reductions and a host-synchronising tolerance test).
//
//   u <- normalize(conv(v; W))        u in R^{Cout*h*w}   (flat CHW)
//   v <- normalize(conv^T(u; W))      v in R^{Cin*h*w}
//   until  ||u-u_old||/sqrt(nu) < atol + rtol*max(u)  and the same for v  (signed max, quirk #8), cap 200
//   sigma = <u, conv(v)>
//
// One side of every shipped layer is narrow (c*h*w = 3072 floats for the CIFAR flows) and one is wide
// (512 channels).  The narrow vector is replicated in every CTA's shared memory; each CTA owns a slice
// of the wide channels (its slice of the wide vector and of the weights stay in shared memory for the
// whole solve).  narrow->wide needs no communication; wide->narrow writes per-CTA partial sums that
// are reduced in a fixed order after a grid barrier (deterministic: no atomics).  3 grid barriers per
// iteration; norms / tolerance statistics ride on the same barriers.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace impflow {

constexpr int kPcThreads = 1024;      // 32 warps: the per-CTA convolutions are shared-memory-latency bound (was 256: 2x slower)
constexpr int kPcWarps = kPcThreads / 32;

struct PcArgs {
  const float* W;   // [Cout][Cin][3][3]
  float* u;
  float* v;
  float* sigma;
  int* iters;
  int Cout, Cin, H, Wd;
  int n_iterations;
  float atol, rtol;
  float* ws;
  int chans;        // wide channels per CTA
  float* D;         // optional [Cout][Cin][3][3]: d sigma / d W at the final (u, v)
};

__device__ __forceinline__ float pc_block_sum(float v, float* scratch) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kPcWarps; ++w) t += scratch[w];
  return t;
}
__device__ __forceinline__ float pc_block_max(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = scratch[0];
#pragma unroll
  for (int w = 1; w < kPcWarps; ++w) t = fmaxf(t, scratch[w]);
  return t;
}
// fixed-order sum / max over the per-CTA partials of the whole grid (identical in every CTA)
__device__ __forceinline__ float pc_grid_sum(const float* part, int G, float* scratch) {
  float a = 0.f;
  for (int i = threadIdx.x; i < G; i += kPcThreads) a += part[i];
  return pc_block_sum(a, scratch);
}
__device__ __forceinline__ float pc_grid_max(const float* part, int G, float* scratch) {
  float a = -INFINITY;
  for (int i = threadIdx.x; i < G; i += kPcThreads) a = fmaxf(a, part[i]);
  return pc_block_max(a, scratch);
}

__global__ void __launch_bounds__(kPcThreads)
k_sn_power_iter_conv3x3(const PcArgs a) {
  cg::grid_group grid = cg::this_grid();
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
  const int H = a.H, Wd = a.Wd, HW = H * Wd;
  const bool wide_is_out = a.Cout >= a.Cin;       // wide = u (conv output side) or v (conv input side)
  const int Cw = wide_is_out ? a.Cout : a.Cin, Cn = wide_is_out ? a.Cin : a.Cout;
  const int Ln = Cn * HW;
  const long long Lw = (long long)Cw * HW;
  const int cw0 = cta * a.chans;
  const int nch = max(0, min(a.chans, Cw - cw0));   // wide channels of this CTA
  const int Ls = nch * HW;                           // its slice of the wide vector
  const int sgn = wide_is_out ? 1 : -1;              // tap direction of narrow->wide

  extern __shared__ float sm[];
  float* nb = sm;                          // narrow vector (normalised)
  float* nb_old = nb + Ln;
  float* wa = nb_old + Ln;                 // wide slice (normalised)
  float* wa_old = wa + a.chans * HW;       // previous wide slice / scratch for raw products
  float* wts = wa_old + a.chans * HW;      // [nch][Cn][9]
  __shared__ float scratch[kPcWarps];

  float* const ws_partial = a.ws;                       // [G][Ln]
  float* const ws_nraw = ws_partial + (size_t)G * Ln;   // [Ln]
  float* const ws_ssq_w = ws_nraw + Ln;                 // [G] each
  float* const ws_err_w = ws_ssq_w + G;
  float* const ws_max_w = ws_err_w + G;
  float* const ws_ssq_n = ws_max_w + G;
  float* const ws_dot = ws_ssq_n + G;

  float* const wide_g = wide_is_out ? a.u : a.v;
  float* const narrow_g = wide_is_out ? a.v : a.u;
  for (int i = tid; i < Ln; i += kPcThreads) nb[i] = narrow_g[i];
  for (int i = tid; i < Ls; i += kPcThreads) wa[i] = wide_g[(long long)cw0 * HW + i];
  for (int i = tid; i < nch * Cn * 9; i += kPcThreads) {
    const int cl = i / (Cn * 9), r = i % (Cn * 9), cn = r / 9, k = r % 9;
    wts[i] = wide_is_out ? a.W[((long long)(cw0 + cl) * a.Cin + cn) * 9 + k]
                         : a.W[((long long)cn * a.Cin + (cw0 + cl)) * 9 + k];
  }
  __syncthreads();

  // narrow -> wide: dst[cl,y,x] = sum_{cn,ky,kx} T(cl,cn,ky,kx) * nb[cn, y + s(ky-1), x + s(kx-1)]
  auto n2w = [&](float* dst) {
    for (int idx = tid; idx < Ls; idx += kPcThreads) {
      const int cl = idx / HW, p = idx % HW, y = p / Wd, x = p % Wd;
      const float* wt = wts + cl * Cn * 9;
      float acc = 0.f;
      for (int cn = 0; cn < Cn; ++cn) {
        const float* src = nb + cn * HW;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int yy = y + sgn * (ky - 1);
          if (yy < 0 || yy >= H) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int xx = x + sgn * (kx - 1);
            if (xx < 0 || xx >= Wd) continue;
            acc += wt[cn * 9 + ky * 3 + kx] * src[yy * Wd + xx];
          }
        }
      }
      dst[idx] = acc;
    }
    __syncthreads();
  };
  // wide -> narrow partial of this CTA's channels: out[cn,y,x] = sum_{cl,ky,kx} T * wa[cl, y - s(ky-1), x - s(kx-1)]
  auto w2n_partial = [&](float* out) {
    for (int i = tid; i < Ln; i += kPcThreads) {
      const int cn = i / HW, p = i % HW, y = p / Wd, x = p % Wd;
      float acc = 0.f;
      for (int cl = 0; cl < nch; ++cl) {
        const float* wt = wts + (cl * Cn + cn) * 9;
        const float* src = wa + cl * HW;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int yy = y - sgn * (ky - 1);
          if (yy < 0 || yy >= H) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int xx = x - sgn * (kx - 1);
            if (xx < 0 || xx >= Wd) continue;
            acc += wt[ky * 3 + kx] * src[yy * Wd + xx];
          }
        }
      }
      out[i] = acc;
    }
  };
  // reduce this CTA's chunk of the narrow vector over all partials (fixed order); optional dot with nb
  const int chunk = (Ln + G - 1) / G;
  const int i0 = min(Ln, cta * chunk), i1 = min(Ln, i0 + chunk);
  auto reduce_chunk = [&](float* ssq_or_dot_out, bool dot_with_nb) {
    const int warp = tid >> 5, lane = tid & 31;
    float local = 0.f;          // accumulated by lane 0 of each warp, in element order
    for (int i = i0 + warp; i < i1; i += kPcWarps) {
      float s = 0.f;
      for (int g = lane; g < G; g += 32) s += ws_partial[(size_t)g * Ln + i];
      s = warp_sum(s);
      if (lane == 0) {
        if (!dot_with_nb) ws_nraw[i] = s;
        local += dot_with_nb ? s * nb[i] : s * s;
      }
    }
    const float tot = pc_block_sum(lane == 0 ? local : 0.f, scratch);
    if (tid == 0) ssq_or_dot_out[cta] = tot;
  };
  // nb <- normalize(nraw); returns (err^2, max) of the narrow vector, identical in every CTA
  auto narrow_update = [&](float& err2, float& mx) {
    const float nrm = fmaxf(sqrtf(pc_grid_sum(ws_ssq_n, G, scratch)), 1e-12f);
    float e = 0.f, m = -INFINITY;
    for (int i = tid; i < Ln; i += kPcThreads) {
      const float old = nb[i];
      const float val = ws_nraw[i] / nrm;
      nb_old[i] = old;
      nb[i] = val;
      e += (val - old) * (val - old);
      m = fmaxf(m, val);
    }
    err2 = pc_block_sum(e, scratch);
    mx = pc_block_max(m, scratch);
    __syncthreads();
  };
  // wa <- wa_raw / norm; publishes this CTA's (err^2, max) partials of the wide vector
  auto wide_update = [&](const float* raw, const float* old) {
    const float nrm = fmaxf(sqrtf(pc_grid_sum(ws_ssq_w, G, scratch)), 1e-12f);
    float e = 0.f, m = -INFINITY;
    for (int i = tid; i < Ls; i += kPcThreads) {
      const float val = raw[i] / nrm;
      const float d = val - old[i];
      e += d * d;
      m = fmaxf(m, val);
      wa[i] = val;
    }
    e = pc_block_sum(e, scratch);
    m = pc_block_max(m, scratch);
    if (tid == 0) {
      ws_err_w[cta] = e;
      ws_max_w[cta] = m;
    }
    __syncthreads();
  };
  auto ssq_of = [&](const float* x, int n, float* out) {
    float s = 0.f;
    for (int i = tid; i < n; i += kPcThreads) s += x[i] * x[i];
    s = pc_block_sum(s, scratch);
    if (tid == 0) out[cta] = s;
  };

  const bool tol_mode = a.n_iterations < 0;
  const int max_it = tol_mode ? 200 : a.n_iterations;
  int it = 0;
  float sigma = 0.f;
  if (wide_is_out) {
    // u wide, v narrow:  u = normalize(n2w(v));  v = normalize(w2n(u))
    while (it < max_it) {
      for (int i = tid; i < Ls; i += kPcThreads) wa_old[i] = wa[i];
      __syncthreads();
      n2w(wa);
      ssq_of(wa, Ls, ws_ssq_w);
      grid.sync();
      wide_update(wa, wa_old);
      w2n_partial(ws_partial + (size_t)cta * Ln);
      grid.sync();
      reduce_chunk(ws_ssq_n, false);
      grid.sync();
      float err_n2, max_n;
      narrow_update(err_n2, max_n);
      ++it;
      if (tol_mode) {
        const float err_u = sqrtf(pc_grid_sum(ws_err_w, G, scratch)) / sqrtf((float)Lw);
        const float tol_u = a.atol + a.rtol * pc_grid_max(ws_max_w, G, scratch);
        const float err_v = sqrtf(err_n2) / sqrtf((float)Ln);
        const float tol_v = a.atol + a.rtol * max_n;
        if (err_u < tol_u && err_v < tol_v) break;
      }
    }
    // sigma = <u, conv(v)> = <wa, n2w(nb)>
    n2w(wa_old);
    float d = 0.f;
    for (int i = tid; i < Ls; i += kPcThreads) d += wa[i] * wa_old[i];
    d = pc_block_sum(d, scratch);
    if (tid == 0) ws_dot[cta] = d;
    grid.sync();
    sigma = pc_grid_sum(ws_dot, G, scratch);
  } else {
    // v wide, u narrow:  u = normalize(w2n(v));  v = normalize(n2w(u))
    bool have_partial = false;
    float err_n2 = 0.f, max_n = 0.f;
    while (it < max_it) {
      w2n_partial(ws_partial + (size_t)cta * Ln);      // conv(v) partials (also what sigma needs)
      grid.sync();
      if (tol_mode && it > 0) {                         // tolerance test of the previous iteration
        const float err_v = sqrtf(pc_grid_sum(ws_err_w, G, scratch)) / sqrtf((float)Lw);
        const float tol_v = a.atol + a.rtol * pc_grid_max(ws_max_w, G, scratch);
        const float err_u = sqrtf(err_n2) / sqrtf((float)Ln);
        const float tol_u = a.atol + a.rtol * max_n;
        if (err_u < tol_u && err_v < tol_v) {
          have_partial = true;
          break;
        }
      }
      reduce_chunk(ws_ssq_n, false);
      grid.sync();
      narrow_update(err_n2, max_n);
      n2w(wa_old);
      ssq_of(wa_old, Ls, ws_ssq_w);
      grid.sync();
      // wa_old holds the raw product, wa the previous v: normalise into wa
      {
        const float nrm = fmaxf(sqrtf(pc_grid_sum(ws_ssq_w, G, scratch)), 1e-12f);
        float e = 0.f, m = -INFINITY;
        for (int i = tid; i < Ls; i += kPcThreads) {
          const float val = wa_old[i] / nrm;
          const float dd = val - wa[i];
          e += dd * dd;
          m = fmaxf(m, val);
          wa[i] = val;
        }
        e = pc_block_sum(e, scratch);
        m = pc_block_max(m, scratch);
        if (tid == 0) {
          ws_err_w[cta] = e;
          ws_max_w[cta] = m;
        }
        __syncthreads();
      }
      ++it;
    }
    if (!have_partial) {
      w2n_partial(ws_partial + (size_t)cta * Ln);
      grid.sync();
    }
    reduce_chunk(ws_dot, true);                         // <u, conv(v)> chunk-wise
    grid.sync();
    sigma = pc_grid_sum(ws_dot, G, scratch);
  }
  if (a.D != nullptr) {
    // d sigma / d W[co][ci][ky][kx] = sum_{y,x} u[co,y,x] v[ci,y+ky-1,x+kx-1]  (sigma = <u, conv(v)> is linear in W):
    // one warp per (wide channel of this CTA, narrow channel, tap), lanes over the pixels
    const int warp = tid >> 5, lane = tid & 31;
    for (int e = warp; e < nch * Cn * 9; e += kPcWarps) {
      const int cl = e / (Cn * 9), r = e % (Cn * 9), cn = r / 9, k = r % 9;
      const int dy = k / 3 - 1, dx = k % 3 - 1;
      const float* wsl = wa + cl * HW;
      const float* nsl = nb + cn * HW;
      float acc = 0.f;
      int y = lane / Wd, x = lane % Wd;          // pixel of this lane, advanced by 32 per step without divisions
      const int step_y = 32 / Wd, step_x = 32 % Wd;
      for (int p = lane; p < HW; p += 32, y += step_y, x += step_x) {
        if (x >= Wd) {
          x -= Wd;
          ++y;
        }
        // u is indexed at the output pixel, v at the shifted input pixel
        const int uy = wide_is_out ? y : y - dy, ux = wide_is_out ? x : x - dx;
        const int vy = wide_is_out ? y + dy : y, vx = wide_is_out ? x + dx : x;
        if (uy < 0 || uy >= H || ux < 0 || ux >= Wd || vy < 0 || vy >= H || vx < 0 || vx >= Wd) continue;
        const float uu = wide_is_out ? wsl[uy * Wd + ux] : nsl[uy * Wd + ux];
        const float vv = wide_is_out ? nsl[vy * Wd + vx] : wsl[vy * Wd + vx];
        acc += uu * vv;
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        const long long co = wide_is_out ? (cw0 + cl) : cn, ci = wide_is_out ? cn : (cw0 + cl);
        a.D[(co * a.Cin + ci) * 9 + k] = acc;
      }
    }
  }
  if (it > 0) {
    for (int i = tid; i < Ls; i += kPcThreads) wide_g[(long long)cw0 * HW + i] = wa[i];
    if (cta == 0)
      for (int i = tid; i < Ln; i += kPcThreads) narrow_g[i] = nb[i];
  }
  if (cta == 0 && tid == 0) {
    a.sigma[0] = sigma;
    if (a.iters != nullptr) a.iters[0] = it;
  }
}

// CTAs per layer.  The solve is bound by its grid barriers, not by arithmetic, and update_lipschitz runs the 24
// conv layers of the CIFAR flow on side streams: with 128 CTAs per launch (the round-1 choice) no two cooperative
// launches fit on the 148 SMs together and the layers serialise (24 x 125 us at the END of every step, on the
// critical path); with 32 CTAs four layers run side by side and each grid barrier is cheaper.
static int g_pc_ctas = 32;

static int pc_plan(int Cout, int Cin, int H, int Wd, int* chans, int* grid, size_t* smem) {
  const int Cw = Cout >= Cin ? Cout : Cin, Cn = Cout >= Cin ? Cin : Cout;
  const long long HW = (long long)H * Wd;
  int c = (Cw + g_pc_ctas - 1) / g_pc_ctas;
  if (c < 1) c = 1;
  // the CTA's slice of the wide vector and of W must fit in shared memory beside the narrow vector
  auto floats = [&](int ch) { return 2 * Cn * HW + 2 * ch * HW + (long long)ch * Cn * 9; };
  while (c > 1 && floats(c) * 4 > 200 * 1024) --c;
  *chans = c;
  *grid = (Cw + c - 1) / c;
  if (*grid > 128) {                        // <= 128 CTAs: co-resident on 148 SMs at one CTA per SM
    c = (Cw + 127) / 128;
    *chans = c;
    *grid = (Cw + c - 1) / c;
  }
  const long long fl = floats(c);
  *smem = (size_t)fl * sizeof(float);
  return (fl * 4 <= 200 * 1024) ? 0 : -1;
}

}  // namespace impflow

using namespace impflow;

extern "C" int impflow_sn_conv_set_ctas(int ctas) {
  const int prev = g_pc_ctas;
  g_pc_ctas = ctas < 1 ? 1 : (ctas > 128 ? 128 : ctas);
  return prev;
}

extern "C" size_t impflow_sn_conv_workspace_floats(int Cout, int Cin, int H, int Wd) {
  int chans, grid;
  size_t smem;
  if (pc_plan(Cout, Cin, H, Wd, &chans, &grid, &smem) != 0) return 0;     // 0: shape not supported
  const size_t Ln = (size_t)(Cout >= Cin ? Cin : Cout) * H * Wd;
  return (size_t)grid * Ln + Ln + 5 * (size_t)grid;
}

extern "C" int impflow_sn_power_iter_conv3x3(const float* W, float* u, float* v, float* sigma, int* iters, int Cout,
                                             int Cin, int H, int Wd, int n_iterations, float atol, float rtol,
                                             float* ws, float* D, void* stream) {
  IMPFLOW_REQUIRE(Cout >= 1 && Cin >= 1 && H >= 1 && Wd >= 1, "sn_power_iter_conv3x3: empty problem");
  int chans, grid;
  size_t smem;
  if (pc_plan(Cout, Cin, H, Wd, &chans, &grid, &smem) != 0) {
    set_error("sn_power_iter_conv3x3: narrow side %d x %d x %d does not fit in shared memory", Cout < Cin ? Cout : Cin,
              H, Wd);
    return -2;
  }
  IMPFLOW_REQUIRE(ws != nullptr, "sn_power_iter_conv3x3: workspace missing");
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    if (cudaFuncSetAttribute(k_sn_power_iter_conv3x3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess) {
      set_error("sn_power_iter_conv3x3: cannot set %zu bytes of dynamic shared memory", smem);
      return -1;
    }
    smem_set = smem;
  }
  PcArgs a{W, u, v, sigma, iters, Cout, Cin, H, Wd, n_iterations, atol, rtol, ws, chans, D};
  void* args[] = {&a};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)k_sn_power_iter_conv3x3, dim3(grid), dim3(kPcThreads), args, smem,
                                              (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_error("sn_power_iter_conv3x3: cooperative launch failed: %s", cudaGetErrorString(e));
    return -1;
  }
  ++g_launch_count;
  return 0;
}
