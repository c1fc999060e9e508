// Elementwise / reduction kernels on the ImpFlow hot path (HBM-bound; coalesced float4 access,
// grids sized as multiples of the 148 SMs).
#include <stdarg.h>

#include "common.cuh"

namespace impflow {

static thread_local char g_err[512] = "";
long long g_launch_count = 0;
int g_pdl = 1;
thread_local const int* g_gate = nullptr;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

constexpr int kSMs = 148;

static inline int grid_for(long long work_items, int per_block) {
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)kSMs * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- act_mul: out = g * act^(order)(x) ---------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256)
k_act_mul(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ out, long long n,
          int order, const float* __restrict__ beta_ptr, int vec) {
  const float beta = (beta_ptr != nullptr) ? __ldg(beta_ptr) : 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (vec) {
    const long long n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (; i < n4; i += stride) {
      const float4 xv = x4[i];
      float4 r;
      r.x = act_eval<KIND>(xv.x, order, beta);
      r.y = act_eval<KIND>(xv.y, order, beta);
      r.z = act_eval<KIND>(xv.z, order, beta);
      r.w = act_eval<KIND>(xv.w, order, beta);
      if (g != nullptr) {
        const float4 gv = g4[i];
        r.x *= gv.x; r.y *= gv.y; r.z *= gv.z; r.w *= gv.w;
      }
      o4[i] = r;
    }
    i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  }
  for (; i < n; i += stride) {
    float r = act_eval<KIND>(x[i], order, beta);
    if (g != nullptr) r *= g[i];
    out[i] = r;
  }
}

// ---- beta gradient of LipSwish: two-stage deterministic reduce --------------------------------
__global__ void __launch_bounds__(256)
k_beta_grad_stage1(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ g2,
                   float* __restrict__ partial, long long n, int order, const float* __restrict__ beta_ptr) {
  const float beta = __ldg(beta_ptr);
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float w = g[i];
    if (g2 != nullptr) w *= g2[i];
    acc += (double)(w * lipswish_dbeta(x[i], order, beta));
  }
  __shared__ double ws[8];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    partial[blockIdx.x] = (float)t;
  }
}
__global__ void k_sum_partials(const float* __restrict__ partial, float* __restrict__ out, int m) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < m; i += 32) acc += (double)partial[i];
  acc = warp_sum_d(acc);
  if (threadIdx.x == 0) out[0] = (float)acc;
}

// ---- second-order activation term of the Neumann-gradient reverse pass:
//      out = act''(p) * t * ga + act'(p) * gb     (gb may be null)
template <int KIND>
__global__ void __launch_bounds__(256)
k_act_second(const float* __restrict__ p, const float* __restrict__ t, const float* __restrict__ ga,
             const float* __restrict__ gb, float* __restrict__ out, long long n, const float* __restrict__ beta_ptr) {
  const float beta = (beta_ptr != nullptr) ? __ldg(beta_ptr) : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    float r = act_eval<KIND>(pv, 2, beta) * t[i] * ga[i];
    if (gb != nullptr) r += act_eval<KIND>(pv, 1, beta) * gb[i];
    out[i] = r;
  }
}

__device__ __forceinline__ void split1(float v, float& h, float& l) {
  uint32_t hb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
  h = __uint_as_float(hb);
  l = v - h;
}
// ---- act^(order)(x) written directly as tf32 hi/lo planes (layer inputs re-evaluated from saved pre-activations
// for the weight gradients: no fp32 round trip, no separate split pass)
template <int KIND>
__global__ void __launch_bounds__(256)
k_act_split(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long n, int order,
            const float* __restrict__ beta_ptr) {
  const float beta = (beta_ptr != nullptr) ? __ldg(beta_ptr) : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float h, l;
    split1(act_eval<KIND>(x[i], order, beta), h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// ---- fused activation step of the Neumann reverse sweep for one LipSwish layer (wide tensors, N columns):
//   ybar = act''(p) * t * ta + act'(p) * ab                (adjoint of the pre-activation; tf32 hi/lo planes)
//   colsum[n]  += sum_rows ybar                           (bias gradient of the layer below)
//   beta_grad  += sum ta * t * d/dbeta act'(p) + ab * d/dbeta act(p)
// One pass over p, t, ta, ab instead of act_second + two act_beta_grad + colsum + split (14 tensor passes -> 6).
// Block b owns a contiguous row chunk and every column: per-block partials, summed by a second stage in
// fixed order (deterministic).
__global__ void __launch_bounds__(256)
k_neumann_act_bwd(const float* __restrict__ p, const float* __restrict__ t, const float* __restrict__ ta,
                  const float* __restrict__ ab, float* __restrict__ y_hi, float* __restrict__ y_lo,
                  float* __restrict__ col_partial, float* __restrict__ beta_partial, long long M, int N,
                  long long rows_per_block, const float* __restrict__ beta_ptr) {
  const float beta = __ldg(beta_ptr);
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  double bacc = 0.0;
  for (int c = threadIdx.x; c < N; c += 256) {
    float cacc = 0.f;
#pragma unroll 4
    for (long long r = r0; r < r1; ++r) {
      const long long i = r * N + c;
      const float pv = p[i], tv = t[i], tav = ta[i];
      const float abv = ab != nullptr ? ab[i] : 0.f;
      const float y = act_eval<IMPFLOW_ACT_LIPSWISH>(pv, 2, beta) * tv * tav +
                      act_eval<IMPFLOW_ACT_LIPSWISH>(pv, 1, beta) * abv;
      float h, l;
      split1(y, h, l);
      y_hi[i] = h;
      y_lo[i] = l;
      cacc += y;
      float bg = tav * tv * lipswish_dbeta(pv, 1, beta);
      if (ab != nullptr) bg += abv * lipswish_dbeta(pv, 0, beta);
      bacc += (double)bg;
    }
    col_partial[(long long)blockIdx.x * N + c] = cacc;
  }
  __shared__ double ws[8];
  bacc = warp_sum_d(bacc);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = bacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tt = 0.0;
    for (int w = 0; w < 8; ++w) tt += ws[w];
    beta_partial[blockIdx.x] = (float)tt;
  }
}
// second stage: colsum[n] = sum_b col_partial[b][n].  One block per 32 columns; 8 row groups sum every 8th partial
// row (coalesced 128-byte reads) and are combined in fixed order: deterministic, and 8x shorter dependent chains
// than one thread per column.
__global__ void __launch_bounds__(256)
k_sum_col_partials(const float* __restrict__ part, float* __restrict__ out, int nblk, int N) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (c < N) {
#pragma unroll 4
    for (int b = grp; b < nblk; b += 8) acc += part[(long long)b * N + c];
  }
  red[grp][lane] = acc;
  __syncthreads();
  if (grp == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += red[g][lane];
    out[c] = t;
  }
}

// ---- lincomb3 -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_lincomb3(const float* __restrict__ a, float ca, const float* __restrict__ b, float cb,
           const float* __restrict__ c, float cc, float* __restrict__ out, long long n, int vec) {
  pdl_trigger();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (vec) {
    const long long n4 = n >> 2;
    for (; i < n4; i += stride) {
      const float4 av = reinterpret_cast<const float4*>(a)[i];
      float4 r = make_float4(av.x * ca, av.y * ca, av.z * ca, av.w * ca);
      if (b != nullptr) {
        const float4 bv = reinterpret_cast<const float4*>(b)[i];
        r.x += bv.x * cb; r.y += bv.y * cb; r.z += bv.z * cb; r.w += bv.w * cb;
      }
      if (c != nullptr) {
        const float4 cv = reinterpret_cast<const float4*>(c)[i];
        r.x += cv.x * cc; r.y += cv.y * cc; r.z += cv.z * cc; r.w += cv.w * cc;
      }
      reinterpret_cast<float4*>(out)[i] = r;
    }
    i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  }
  for (; i < n; i += stride) {
    float r = a[i] * ca;
    if (b != nullptr) r += b[i] * cb;
    if (c != nullptr) r += c[i] * cc;
    out[i] = r;
  }
}

// ---- rowdot: one CTA per sample (large d) or one warp per sample (small d) -----------------------
__global__ void __launch_bounds__(256)
k_rowdot_block(const float* __restrict__ a, const float* __restrict__ c, float* __restrict__ out, long long d,
               float alpha, float beta, int vec) {
  const long long base = (long long)blockIdx.x * d;
  float acc = 0.f;
  if (vec) {
    const float4* a4 = reinterpret_cast<const float4*>(a + base);
    const float4* c4 = reinterpret_cast<const float4*>(c + base);
    for (long long i = threadIdx.x; i < (d >> 2); i += 256) {
      const float4 x = a4[i], y = c4[i];
      acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
    }
  } else {
    for (long long i = threadIdx.x; i < d; i += 256) acc += a[base + i] * c[base + i];
  }
  __shared__ float ws[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += ws[w];
    out[blockIdx.x] = (beta == 0.f ? 0.f : beta * out[blockIdx.x]) + alpha * t;
  }
}
__global__ void __launch_bounds__(256)
k_rowdot_warp(const float* __restrict__ a, const float* __restrict__ c, float* __restrict__ out, int B, int d,
              float alpha, float beta) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int i = lane; i < d; i += 32) acc += a[(long long)b * d + i] * c[(long long)b * d + i];
  acc = warp_sum(acc);
  if (lane == 0) out[b] = (beta == 0.f ? 0.f : beta * out[b]) + alpha * acc;
}

// ---- colsum: out[n] = sum_m a[m,n]; grid (column tiles of 32, row chunks), two deterministic stages ----
__global__ void __launch_bounds__(256)
k_colsum(const float* __restrict__ a, float* __restrict__ out, long long M, int N, long long rows_per_chunk) {
  __shared__ float tile[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  const long long m0 = (long long)blockIdx.y * rows_per_chunk;
  const long long m1 = m0 + rows_per_chunk < M ? m0 + rows_per_chunk : M;
  float acc = 0.f;
  if (col < N)
    for (long long m = m0 + ry; m < m1; m += 8) acc += a[m * N + col];
  tile[ry][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ry == 0 && col < N) {
    float t = 0.f;
    for (int r = 0; r < 8; ++r) t += tile[r][threadIdx.x & 31];
    out[(long long)blockIdx.y * N + col] = t;
  }
}

// ---- transpose ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_transpose(const float* __restrict__ a, float* __restrict__ out, long long M, long long N) {
  __shared__ float tile[32][33];
  const long long m0 = (long long)blockIdx.y * 32, n0 = (long long)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (m0 + r < M && n0 + tx < N) tile[r][tx] = a[(m0 + r) * N + n0 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (n0 + r < N && m0 + tx < M) out[(n0 + r) * M + m0 + tx] = tile[tx][r];
}

// ---- im2col / col2im for NHWC 3x3 stride 1 pad 1 -------------------------------------------------
// col row p=(b,y,x) has 9*C entries ordered (ky,kx,c): see k_im2col3x3_rows below.

// x[p,c] = sum_tap col[p - off(tap), tap, c] followed by the fused epilogue.
__global__ void __launch_bounds__(256)
k_col2im3x3(const float* __restrict__ col, int B, int H, int W, int C, Epilogue ep_in) {
  pdl_trigger();
  pdl_wait();
  const Epilogue ep = resolve_beta(ep_in);
  const long long total = (long long)B * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = i / C;
    const int xx = (int)(p % W);
    const int yy = (int)((p / W) % H);
    const long long b = p / ((long long)W * H);
    float acc = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int sy = yy - (tap / 3 - 1), sx = xx - (tap % 3 - 1);
      if (sy >= 0 && sy < H && sx >= 0 && sx < W)
        acc += col[(((b * H + sy) * W + sx) * 9 + tap) * C + c];
    }
    epilogue_store(ep, p, c, acc);
  }
}

__global__ void __launch_bounds__(256)
k_split_tf32(const float* __restrict__ a, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = a[i];
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    const float hf = __uint_as_float(h);
    hi[i] = hf;
    lo[i] = v - hf;
  }
}
__global__ void __launch_bounds__(256)
k_split_tf32_v4(const float4* __restrict__ a, float4* __restrict__ hi, float4* __restrict__ lo, long long n4) {
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = a[i];
    float4 h, l;
    split1(v.x, h.x, l.x);
    split1(v.y, h.y, l.y);
    split1(v.z, h.z, l.z);
    split1(v.w, h.w, l.w);
    hi[i] = h;
    lo[i] = l;
  }
}

// out_hi/out_lo[n,m] = tf32 split of a[m,n]: the transposed operand planes of the weight-gradient GEMMs
// in one pass (instead of transpose, then split).  64 x 32 tiles, 128-byte segments on both sides.
__global__ void __launch_bounds__(256)
k_transpose_split(const float* __restrict__ a, float* __restrict__ out_hi, float* __restrict__ out_lo, long long M,
                  int N) {
  __shared__ float tile[64][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long m_tiles = (M + 63) / 64;
  const int n_tiles = (N + 31) / 32;
  for (long long t = blockIdx.x; t < m_tiles * n_tiles; t += gridDim.x) {
    const long long m0 = (t / n_tiles) * 64;
    const int n0 = (int)(t % n_tiles) * 32;
    for (int r = ty; r < 64; r += 8)
      tile[r][tx] = (m0 + r < M && n0 + tx < N) ? a[(m0 + r) * N + n0 + tx] : 0.f;
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      if (n0 + r < N) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const long long m = m0 + half * 32 + tx;
          if (m < M) {
            float h, l;
            split1(tile[half * 32 + tx][r], h, l);
            out_hi[(long long)(n0 + r) * M + m] = h;
            out_lo[(long long)(n0 + r) * M + m] = l;
          }
        }
      }
    }
    __syncthreads();
  }
}

// im2col: one thread per 4 consecutive patch columns of one pixel (ld % 4 == 0) or per column (otherwise):
// the stores of a warp are one contiguous 512-byte run, the gathers hit the (small, cached) image; optional
// tf32 hi/lo planes instead of fp32.
template <int VEC>
__global__ void __launch_bounds__(256)
k_im2col3x3_rows(const float* __restrict__ x, float* __restrict__ col, float* __restrict__ col_lo, int B, int H,
                 int W, int C, int ld) {
  pdl_trigger();
  pdl_wait();
  const int groups = ld / VEC;
  const long long total = (long long)B * H * W * groups;
  const int K = 9 * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / groups;
    const int k0 = (int)(i - p * groups) * VEC;
    const int xx = (int)(p % W);
    const int yy = (int)((p / W) % H);
    const long long img = p - (long long)yy * W - xx;          // pixel index of (b, 0, 0)
    float v[VEC];
#pragma unroll
    for (int u = 0; u < VEC; ++u) {
      const int k = k0 + u;
      v[u] = 0.f;
      if (k < K) {
        const int tap = k / C, c = k - tap * C;
        const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
        if (sy >= 0 && sy < H && sx >= 0 && sx < W) v[u] = __ldg(x + (img + (long long)sy * W + sx) * C + c);
      }
    }
    float* dst = col + p * ld + k0;
    if (col_lo != nullptr) {
      float h[VEC], l[VEC];
#pragma unroll
      for (int u = 0; u < VEC; ++u) split1(v[u], h[u], l[u]);
      float* dst_lo = col_lo + p * ld + k0;
      if (VEC == 4) {
        *reinterpret_cast<float4*>(dst) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(dst_lo) = make_float4(l[0], l[1], l[2], l[3]);
      } else {
        dst[0] = h[0];
        dst_lo[0] = l[0];
      }
    } else if (VEC == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      dst[0] = v[0];
    }
  }
}

// Step tail in one pass over the flat parameter / gradient buffers (train_img.py:652-658): gradient clipping
// (clip_grad_norm_: g *= min(1, max_norm / (||g|| + 1e-6)), the norm read from the device), the vendored Adam
// update (lib/optimizers.py:86-103: denom = sqrt(v) + eps, step = lr * sqrt(1-b2^t) / (1-b1^t)) and the
// exponential moving average of the parameters (lib/utils.py:140-146).
__global__ void __launch_bounds__(256)
k_clip_adam_ema(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                float* __restrict__ ema, long long n, const float* __restrict__ gnorm_sq, float max_norm,
                float step_size, float beta1, float beta2, float eps, float ema_decay) {
  float coef = 1.f;
  if (gnorm_sq != nullptr && max_norm > 0.f) coef = fminf(1.f, max_norm / (sqrtf(__ldg(gnorm_sq)) + 1e-6f));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    const float pi = p[i] - step_size * mi / (sqrtf(vi) + eps);
    g[i] = gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
    if (ema != nullptr) {      // shadow -= (1 - decay) * (shadow - param)   (lib/utils.py:143-146)
      const float e = ema[i];
      ema[i] = e - (1.f - ema_decay) * (e - pi);
    }
  }
}

}  // namespace impflow

namespace impflow {

// ---- ActNorm (act_norm.py:39-62): y = (x + bias_c) exp(w_c), logpx' = logpx - HW sum_c w_c ------------------------
// x is (B, C, HW) contiguous (HW = 1 for ActNorm1d) or, channels_last, (B, HW, C) — the memory order the branch
// kernels leave their outputs in.  One launch forward; the backward is a per-(channel, slice)
// pass that writes gx and the partial sums of gbias / gweight, and a per-channel finish in fixed order.
constexpr int kActNormSlices = 32;

__global__ void __launch_bounds__(256)
k_actnorm_fwd(const float* __restrict__ x, const float* __restrict__ bias, const float* __restrict__ w,
              float* __restrict__ y, const float* __restrict__ logpx, float* __restrict__ logpx_out, long long B, int C,
              long long HW, int channels_last) {
  const long long n = B * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = channels_last ? (int)(i % C) : (int)((i / HW) % C);
    y[i] = (x[i] + __ldg(bias + c)) * expf(__ldg(w + c));
  }
  if (logpx_out != nullptr && blockIdx.x == 0) {
    __shared__ float ld;
    if (threadIdx.x == 0) {
      float sacc = 0.f;
      for (int c = 0; c < C; ++c) sacc += w[c];          // fixed order
      ld = sacc * (float)HW;
    }
    __syncthreads();
    for (long long b = threadIdx.x; b < B; b += blockDim.x) logpx_out[b] = logpx[b] - ld;
  }
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  red[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  const float r = red[0];
  __syncthreads();
  return r;
}

// grid (kActNormSlices, C): slice of the B*HW positions of channel c
__global__ void __launch_bounds__(256)
k_actnorm_bwd_partial(const float* __restrict__ gy, const float* __restrict__ y, const float* __restrict__ w,
                      float* __restrict__ gx, float* __restrict__ partial, long long B, int C, long long HW,
                      int channels_last) {
  __shared__ float red[256];
  const int c = blockIdx.y;
  const float e = expf(__ldg(w + c));
  const long long per = B * HW;
  const long long chunk = (per + kActNormSlices - 1) / kActNormSlices;
  const long long j0 = blockIdx.x * chunk, j1 = min(per, j0 + chunk);
  float sb = 0.f, sw = 0.f;
  for (long long j = j0 + threadIdx.x; j < j1; j += 256) {
    const long long i = channels_last ? j * C + c : ((j / HW) * C + c) * HW + (j % HW);
    const float g = gy[i];
    gx[i] = g * e;
    sb = fmaf(g, e, sb);
    sw = fmaf(g, y[i], sw);
  }
  sb = block_sum_256(sb, red);
  sw = block_sum_256(sw, red);
  if (threadIdx.x == 0) {
    partial[((long long)c * kActNormSlices + blockIdx.x) * 2] = sb;
    partial[((long long)c * kActNormSlices + blockIdx.x) * 2 + 1] = sw;
  }
}

// one block per channel: gbias_c, gweight_c = sum gy*y - HW * sum_b g_logpx[b]
__global__ void __launch_bounds__(256)
k_actnorm_bwd_finish(const float* __restrict__ partial, const float* __restrict__ g_logpx, float* __restrict__ gbias,
                     float* __restrict__ gw, long long B, long long HW) {
  __shared__ float red[256];
  const int c = blockIdx.x;
  float gl = 0.f;
  if (g_logpx != nullptr)
    for (long long b = threadIdx.x; b < B; b += 256) gl += g_logpx[b];
  gl = block_sum_256(gl, red);
  if (threadIdx.x == 0) {
    float sb = 0.f, sw = 0.f;
    for (int s = 0; s < kActNormSlices; ++s) {
      sb += partial[((long long)c * kActNormSlices + s) * 2];
      sw += partial[((long long)c * kActNormSlices + s) * 2 + 1];
    }
    gbias[c] = sb;
    gw[c] = sw - (float)HW * gl;
  }
}

}  // namespace impflow

using namespace impflow;

extern "C" int impflow_clip_adam_ema(float* p, float* g, float* m, float* v, float* ema, long long n,
                                     const float* gnorm_sq, float max_norm, float step_size, float beta1,
                                     float beta2, float eps, float ema_decay, void* stream) {
  if (n <= 0) return 0;
  k_clip_adam_ema<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, ema, n, gnorm_sq, max_norm, step_size,
                                                                     beta1, beta2, eps, ema_decay);
  return check_launch("k_clip_adam_ema");
}

extern "C" int impflow_actnorm_forward(const float* x, const float* bias, const float* weight, float* y,
                                       const float* logpx, float* logpx_out, long long B, int C, long long HW,
                                       int channels_last, void* stream) {
  IMPFLOW_REQUIRE(B >= 1 && C >= 1 && HW >= 1, "actnorm_forward: empty input");
  IMPFLOW_REQUIRE((logpx == nullptr) == (logpx_out == nullptr), "actnorm_forward: logpx and logpx_out come together");
  k_actnorm_fwd<<<grid_for(B * C * HW, 256), 256, 0, (cudaStream_t)stream>>>(x, bias, weight, y, logpx, logpx_out, B, C,
                                                                             HW, channels_last);
  return check_launch("k_actnorm_fwd");
}

extern "C" size_t impflow_actnorm_workspace_floats(int C) { return (size_t)C * kActNormSlices * 2; }

extern "C" int impflow_actnorm_backward(const float* gy, const float* y, const float* weight, const float* g_logpx,
                                        float* gx, float* gbias, float* gweight, float* ws, long long B, int C,
                                        long long HW, int channels_last, void* stream) {
  IMPFLOW_REQUIRE(B >= 1 && C >= 1 && HW >= 1, "actnorm_backward: empty input");
  IMPFLOW_REQUIRE(C <= 65535, "actnorm_backward: C=%d too large", C);
  IMPFLOW_REQUIRE(ws != nullptr, "actnorm_backward: workspace missing");
  cudaStream_t s = (cudaStream_t)stream;
  k_actnorm_bwd_partial<<<dim3(kActNormSlices, C), 256, 0, s>>>(gy, y, weight, gx, ws, B, C, HW, channels_last);
  if (check_launch("k_actnorm_bwd_partial")) return -1;
  k_actnorm_bwd_finish<<<C, 256, 0, s>>>(ws, g_logpx, gbias, gweight, B, HW);
  return check_launch("k_actnorm_bwd_finish");
}

extern "C" int impflow_version(void) { return IMPFLOW_ABI_VERSION; }
extern "C" const char* impflow_last_error(void) { return impflow::g_err; }
extern "C" long long impflow_launch_count(void) { return impflow::g_launch_count; }
extern "C" void impflow_add_launch_count(long long n) { impflow::g_launch_count += n; }
extern "C" int impflow_set_pdl(int on) {
  const int prev = impflow::g_pdl;
  impflow::g_pdl = on ? 1 : 0;
  return prev;
}

extern "C" int impflow_act_mul(const float* x, const float* g, float* out, long long n, int kind, int order,
                               const float* beta_sp, void* stream) {
  IMPFLOW_REQUIRE(kind != IMPFLOW_ACT_LIPSWISH || beta_sp != nullptr, "act_mul: LipSwish needs beta_sp");
  IMPFLOW_REQUIRE(order >= 0 && order <= 3, "act_mul: order %d not in [0,3]", order);
  if (n <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int vec = (aligned16(x) && aligned16(out) && (g == nullptr || aligned16(g)) && n >= 4) ? 1 : 0;
  const int grid = grid_for(vec ? (n >> 2) : n, 256);
  switch (kind) {
    case IMPFLOW_ACT_SIN: k_act_mul<IMPFLOW_ACT_SIN><<<grid, 256, 0, s>>>(x, g, out, n, order, beta_sp, vec); break;
    case IMPFLOW_ACT_LIPSWISH:
      k_act_mul<IMPFLOW_ACT_LIPSWISH><<<grid, 256, 0, s>>>(x, g, out, n, order, beta_sp, vec);
      break;
    case IMPFLOW_ACT_RELU: k_act_mul<IMPFLOW_ACT_RELU><<<grid, 256, 0, s>>>(x, g, out, n, order, beta_sp, vec); break;
    case IMPFLOW_ACT_NONE: k_act_mul<IMPFLOW_ACT_NONE><<<grid, 256, 0, s>>>(x, g, out, n, order, beta_sp, vec); break;
    default: set_error("act_mul: unknown activation kind %d", kind); return -3;
  }
  return check_launch("k_act_mul");
}

extern "C" size_t impflow_reduce_workspace_floats(long long n) {
  (void)n;
  return (size_t)kSMs * 16;
}

extern "C" int impflow_act_second(const float* p, const float* t, const float* ga, const float* gb, float* out,
                                  long long n, int kind, const float* beta_sp, void* stream) {
  if (n <= 0) return 0;
  IMPFLOW_REQUIRE(kind != IMPFLOW_ACT_LIPSWISH || beta_sp != nullptr, "act_second: LipSwish needs beta_sp");
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = grid_for(n, 256);
  switch (kind) {
    case IMPFLOW_ACT_SIN: k_act_second<IMPFLOW_ACT_SIN><<<grid, 256, 0, s>>>(p, t, ga, gb, out, n, beta_sp); break;
    case IMPFLOW_ACT_LIPSWISH: k_act_second<IMPFLOW_ACT_LIPSWISH><<<grid, 256, 0, s>>>(p, t, ga, gb, out, n, beta_sp); break;
    case IMPFLOW_ACT_RELU: k_act_second<IMPFLOW_ACT_RELU><<<grid, 256, 0, s>>>(p, t, ga, gb, out, n, beta_sp); break;
    case IMPFLOW_ACT_NONE: k_act_second<IMPFLOW_ACT_NONE><<<grid, 256, 0, s>>>(p, t, ga, gb, out, n, beta_sp); break;
    default: set_error("act_second: unknown activation kind %d", kind); return -3;
  }
  return check_launch("k_act_second");
}

extern "C" int impflow_act_beta_grad(const float* x, const float* g, const float* g2, float* out, float* partial,
                                     long long n, int order, const float* beta_sp, void* stream) {
  IMPFLOW_REQUIRE(beta_sp != nullptr, "act_beta_grad: beta_sp is null");
  IMPFLOW_REQUIRE(order >= 0 && order <= 2, "act_beta_grad: order %d not in [0,2]", order);
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = grid_for(n, 1024);
  k_beta_grad_stage1<<<grid, 256, 0, s>>>(x, g, g2, partial, n, order, beta_sp);
  if (check_launch("k_beta_grad_stage1")) return -1;
  k_sum_partials<<<1, 32, 0, s>>>(partial, out, grid);
  return check_launch("k_sum_partials");
}

extern "C" int impflow_act_split(const float* x, float* hi, float* lo, long long n, int kind, int order,
                                 const float* beta_sp, void* stream) {
  if (n <= 0) return 0;
  IMPFLOW_REQUIRE(order >= 0 && order <= 3, "act_split: order %d not in [0,3]", order);
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = grid_for(n, 256);
  switch (kind) {
    case IMPFLOW_ACT_SIN: k_act_split<IMPFLOW_ACT_SIN><<<grid, 256, 0, s>>>(x, hi, lo, n, order, beta_sp); break;
    case IMPFLOW_ACT_LIPSWISH: k_act_split<IMPFLOW_ACT_LIPSWISH><<<grid, 256, 0, s>>>(x, hi, lo, n, order, beta_sp); break;
    case IMPFLOW_ACT_RELU: k_act_split<IMPFLOW_ACT_RELU><<<grid, 256, 0, s>>>(x, hi, lo, n, order, beta_sp); break;
    case IMPFLOW_ACT_NONE: k_act_split<IMPFLOW_ACT_NONE><<<grid, 256, 0, s>>>(x, hi, lo, n, order, beta_sp); break;
    default: set_error("act_split: unknown activation kind %d", kind); return -3;
  }
  return check_launch("k_act_split");
}

static int neumann_blocks(long long M) {
  long long b = (M + 15) / 16;
  if (b > 148 * 4) b = 148 * 4;
  return b < 1 ? 1 : (int)b;
}

extern "C" size_t impflow_neumann_act_bwd_workspace_floats(long long M, int N) {
  return (size_t)neumann_blocks(M) * ((size_t)N + 1);
}

extern "C" int impflow_neumann_act_bwd(const float* p, const float* t, const float* ta, const float* ab, float* y_hi,
                                       float* y_lo, float* colsum, float* beta_grad, float* ws, long long M, int N,
                                       const float* beta_sp, void* stream) {
  IMPFLOW_REQUIRE(M >= 1 && N >= 1, "neumann_act_bwd: empty problem");
  IMPFLOW_REQUIRE(beta_sp != nullptr && ws != nullptr, "neumann_act_bwd: beta / workspace missing");
  cudaStream_t s = (cudaStream_t)stream;
  const int nblk = neumann_blocks(M);
  const long long rows = (M + nblk - 1) / nblk;
  float* col_partial = ws;
  float* beta_partial = ws + (size_t)nblk * N;
  k_neumann_act_bwd<<<nblk, 256, 0, s>>>(p, t, ta, ab, y_hi, y_lo, col_partial, beta_partial, M, N, rows, beta_sp);
  if (check_launch("k_neumann_act_bwd")) return -1;
  k_sum_col_partials<<<(N + 31) / 32, 256, 0, s>>>(col_partial, colsum, nblk, N);
  if (check_launch("k_sum_col_partials")) return -1;
  k_sum_partials<<<1, 32, 0, s>>>(beta_partial, beta_grad, nblk);
  return check_launch("k_sum_partials");
}

extern "C" int impflow_lincomb3(const float* a, float ca, const float* b, float cb, const float* c, float cc,
                                float* out, long long n, void* stream) {
  if (n <= 0) return 0;
  const int vec = (aligned16(a) && aligned16(out) && (b == nullptr || aligned16(b)) &&
                   (c == nullptr || aligned16(c)) && n >= 4) ? 1 : 0;
  const int grid = grid_for(vec ? (n >> 2) : n, 256);
  k_lincomb3<<<grid, 256, 0, (cudaStream_t)stream>>>(a, ca, b, cb, c, cc, out, n, vec);
  return check_launch("k_lincomb3");
}

extern "C" int impflow_rowdot(const float* a, const float* c, float* out, int B, long long d, float alpha,
                              float beta, void* stream) {
  if (B <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (d <= 256) {
    k_rowdot_warp<<<(B + 7) / 8, 256, 0, s>>>(a, c, out, B, (int)d, alpha, beta);
    return check_launch("k_rowdot_warp");
  }
  const int vec = (aligned16(a) && aligned16(c) && (d % 4 == 0)) ? 1 : 0;
  k_rowdot_block<<<B, 256, 0, s>>>(a, c, out, d, alpha, beta, vec);
  return check_launch("k_rowdot_block");
}

extern "C" int impflow_colsum_chunks(long long M, int N) {
  const int col_tiles = (N + 31) / 32;
  long long chunks = (296 + col_tiles - 1) / col_tiles;
  if (chunks > (M + 63) / 64) chunks = (M + 63) / 64;
  return chunks < 1 ? 1 : (int)chunks;
}

extern "C" int impflow_colsum(const float* a, float* out, float* partial, long long M, int N, void* stream) {
  if (N <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int chunks = impflow_colsum_chunks(M, N);
  const long long rows = (M + chunks - 1) / chunks;
  dim3 grid((N + 31) / 32, chunks);
  if (chunks == 1) {
    k_colsum<<<grid, 256, 0, s>>>(a, out, M, N, rows);
    return check_launch("k_colsum");
  }
  IMPFLOW_REQUIRE(partial != nullptr, "colsum: needs a partial workspace of chunks*N floats");
  k_colsum<<<grid, 256, 0, s>>>(a, partial, M, N, rows);
  if (check_launch("k_colsum")) return -1;
  k_colsum<<<dim3((N + 31) / 32, 1), 256, 0, s>>>(partial, out, chunks, N, chunks);
  return check_launch("k_colsum(stage2)");
}

extern "C" int impflow_transpose(const float* a, float* out, long long M, long long N, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  dim3 grid((unsigned)((N + 31) / 32), (unsigned)((M + 31) / 32));
  IMPFLOW_REQUIRE(grid.y <= 65535, "transpose: M=%lld too large for this grid layout", M);
  k_transpose<<<grid, 256, 0, (cudaStream_t)stream>>>(a, out, M, N);
  return check_launch("k_transpose");
}

static int launch_im2col(const float* x, float* col, float* col_lo, int B, int H, int W, int C, int ld,
                         void* stream) {
  IMPFLOW_REQUIRE(ld >= 9 * C, "im2col3x3: ld=%d < 9*C=%d", ld, 9 * C);
  const long long n_pix = (long long)B * H * W;
  if (n_pix <= 0) return 0;
  const bool vec = (ld % 4 == 0) && aligned16(col) && (col_lo == nullptr || aligned16(col_lo));
  if (vec) {
    k_im2col3x3_rows<4><<<grid_for(n_pix * (ld / 4), 256), 256, 0, (cudaStream_t)stream>>>(x, col, col_lo, B, H, W, C,
                                                                                           ld);
  } else {
    k_im2col3x3_rows<1><<<grid_for(n_pix * ld, 256), 256, 0, (cudaStream_t)stream>>>(x, col, col_lo, B, H, W, C, ld);
  }
  return check_launch("k_im2col3x3_rows");
}

extern "C" int impflow_im2col3x3(const float* x, float* col, int B, int H, int W, int C, int ld, void* stream) {
  return launch_im2col(x, col, nullptr, B, H, W, C, ld, stream);
}

extern "C" int impflow_im2col3x3_split(const float* x, float* col_hi, float* col_lo, int B, int H, int W, int C,
                                       int ld, void* stream) {
  IMPFLOW_REQUIRE(col_hi != nullptr && col_lo != nullptr, "im2col3x3_split: both planes are required");
  return launch_im2col(x, col_hi, col_lo, B, H, W, C, ld, stream);
}

extern "C" int impflow_transpose_split(const float* a, float* out_hi, float* out_lo, long long M, long long N,
                                       void* stream) {
  if (M <= 0 || N <= 0) return 0;
  IMPFLOW_REQUIRE(N < (1LL << 31), "transpose_split: N=%lld too large", N);
  const long long tiles = ((M + 63) / 64) * ((N + 31) / 32);
  k_transpose_split<<<grid_for(tiles, 1), 256, 0, (cudaStream_t)stream>>>(a, out_hi, out_lo, M, (int)N);
  return check_launch("k_transpose_split");
}

extern "C" int impflow_col2im3x3(const float* col, int B, int H, int W, int C, const float* bias, float* pre_out,
                                 float* act_out, const float* dmul_pre, int act_kind, const float* beta_sp,
                                 void* stream) {
  const long long total = (long long)B * H * W * C;
  if (total <= 0) return 0;
  IMPFLOW_REQUIRE(pre_out != nullptr || act_out != nullptr, "col2im3x3: no output given");
  IMPFLOW_REQUIRE(dmul_pre == nullptr || pre_out != nullptr || act_out != nullptr,
                  "col2im3x3: dmul_pre needs pre_out or act_out");
  Epilogue ep{bias, pre_out, act_out, dmul_pre, (long long)C, act_kind, beta_sp, 0.f};
  k_col2im3x3<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(col, B, H, W, C, ep);
  return check_launch("k_col2im3x3");
}

extern "C" int impflow_split_tf32(const float* a, float* hi, float* lo, long long n, void* stream) {
  if (n <= 0) return 0;
  if (aligned16(a) && aligned16(hi) && aligned16(lo) && (n % 4 == 0)) {
    k_split_tf32_v4<<<grid_for(n >> 2, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(a), reinterpret_cast<float4*>(hi), reinterpret_cast<float4*>(lo), n >> 2);
    return check_launch("k_split_tf32_v4");
  }
  k_split_tf32<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, hi, lo, n);
  return check_launch("k_split_tf32");
}
