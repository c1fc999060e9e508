// Broyden solver algebra on the device — replaces lib/layers/broyden.py:101-193.
//
// Data layout in HBM (all fp32, dense):
//   x, g, xn, gn, low_x, low_g : (B, d)
//   Ut, Vt (history)           : (B, T, d)   rank index j contiguous over d
//   sample_sq, low_sq          : (B)
// One solver iteration = one `norm_decide` launch (per-sample ||gn||^2, batch-global norm,
// bookkeeping and break rules in the elected last block) + one `update` launch (rank-1
// update and next iterate).  The update kernel gives each sample to a thread-block CLUSTER:
// every CTA owns a contiguous slice of d, slices exchange their partial dot products through
// distributed shared memory, so no global atomics and a fixed summation order.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace impflow {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxT = 63;

// ------------------------------------------------------------------------------------------
// per-sample ||g||^2 partials + last-block bookkeeping
// ------------------------------------------------------------------------------------------
__device__ void decide(impflow_broyden_state* st, double total, bool init) {
  const double obj = (double)(float)sqrt(total);  // torch.norm(...).item(): fp32 value as a python float
  st->objective = obj;
  if (init) {  // broyden.py:136-151
    st->nstep = 0;
    st->lowest_step = 0;
    st->init_objective = obj;
    st->lowest = obj;
    st->trace[0] = obj;
    st->prot_break = 0;
    st->converged = 0;
    st->stagnated = 0;
    st->do_update = 0;
    st->new_lowest = 0;
    st->active = (obj >= st->eps && 0 < st->threshold) ? 1 : 0;
    return;
  }
  const int T = st->threshold;
  const int nstep = st->nstep + 1;  // :155
  st->nstep = nstep;
  st->trace[nstep] = obj;           // :158
  int new_low = 0;
  if (obj < st->lowest) {           // :159-162
    st->lowest = obj;
    st->lowest_step = nstep;
    new_low = 1;
  }
  st->new_lowest = new_low;
  const int conv = obj < st->eps ? 1 : 0;  // :163
  int stag = 0;
  if (!conv && obj < 3.0 * st->eps && nstep == T) {  // :165-168, trace[-T:]
    double mx = st->trace[nstep - T + 1], mn = mx;
    for (int i = nstep - T + 1; i <= nstep; ++i) {
      mx = fmax(mx, st->trace[i]);
      mn = fmin(mn, st->trace[i]);
    }
    stag = (mx / mn < 1.3) ? 1 : 0;
  }
  int prot = 0;
  if (!conv && !stag && obj > st->init_objective * 1e6) prot = 1;  // :169-172
  st->converged = conv;
  st->stagnated = stag;
  st->prot_break = prot;
  const int upd = (conv || stag || prot) ? 0 : 1;
  st->do_update = upd;
  st->active = (upd && obj >= st->eps && nstep < T) ? 1 : 0;  // while-condition :153
}

// grid (S, B): CTA (s,b) reduces slice s of sample b.
__global__ void __launch_bounds__(kThreads)
k_norm_decide(const float* __restrict__ g, float* __restrict__ partial, float* __restrict__ sample_sq,
              float* __restrict__ low_sq, impflow_broyden_state* st, int B, long long d, int S, int init, int gated,
              BroydenProgress* progress) {
  // speculative iteration of the sync-free loop: the while-condition already failed on the device (uniform)
  if (gated && ((volatile impflow_broyden_state*)st)->active == 0) return;
  const int b = blockIdx.y, s = blockIdx.x;
  const long long chunk = (d + S - 1) / S;
  const long long lo = (long long)s * chunk;
  const long long hi = lo + chunk < d ? lo + chunk : d;
  const float* gp = g + (long long)b * d;
  float acc = 0.f;
  if (((d & 3) == 0) && ((chunk & 3) == 0)) {
    const float4* g4 = reinterpret_cast<const float4*>(gp);
    for (long long i = (lo >> 2) + threadIdx.x; i < (hi >> 2); i += kThreads) {
      const float4 v = g4[i];
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += kThreads) acc += gp[i] * gp[i];
  }
  __shared__ float wsum[kWarps];
  __shared__ double dsum[kWarps];
  __shared__ int is_last;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kWarps; ++w) t += wsum[w];
    partial[(long long)b * S + s] = t;
    __threadfence();
    const int done = atomicAdd(&st->counter, 1);
    is_last = (done == S * B - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // elected block: fixed-order finalisation (deterministic)
  double tot = 0.0;
  for (int i = threadIdx.x; i < B; i += kThreads) {
    float t = 0.f;
    for (int k = 0; k < S; ++k) t += __ldcg(&partial[(long long)i * S + k]);
    sample_sq[i] = t;
    tot += (double)t;
  }
  tot = warp_sum_d(tot);
  if ((threadIdx.x & 31) == 0) dsum[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kWarps; ++w) t += dsum[w];
    st->counter = 0;
    decide(st, t, init != 0);
    __threadfence();
    if (progress != nullptr) {      // mapped pinned host memory: the host polls this record instead of syncing
      volatile BroydenProgress* pr = progress + st->nstep;
      pr->active = st->active;
      __threadfence_system();
      pr->seq = st->nstep;
      __threadfence_system();
    }
  }
  __syncthreads();
  if (init || ((volatile impflow_broyden_state*)st)->new_lowest) {
    for (int i = threadIdx.x; i < B; i += kThreads) low_sq[i] = sample_sq[i];
  }
}

// ------------------------------------------------------------------------------------------
// begin: low_x = x0, low_g = g0, xn = x0 - g0
// ------------------------------------------------------------------------------------------
__global__ void k_begin(const float* __restrict__ x0, const float* __restrict__ g0, float* __restrict__ xn,
                        float* __restrict__ low_x, float* __restrict__ low_g, long long n,
                        impflow_broyden_state* st, int threshold, double eps) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->threshold = threshold;
    st->eps = eps;
    st->counter = 0;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float x = x0[i], g = g0[i];
    low_x[i] = x;
    low_g[i] = g;
    xn[i] = x + (-g);
  }
}

// ------------------------------------------------------------------------------------------
// rank-1 update, cluster per sample.  VEC = 4 (float4 path, d % 4 == 0) or 1.
// dynamic smem: sdx[SL] sdg[SL] sgn[SL] | wsum[3*T][kWarps] | part0[3*T] part1[2] | tot[3*T+2]
// ------------------------------------------------------------------------------------------
template <int LD>
__device__ __forceinline__ float4 ld_hist(const float4* p) {
  return LD == 0 ? __ldg(p) : __ldcg(p);
}

// NT threads per CTA (256: two CTAs of 8 warps per SM; 512: two CTAs of 16 warps), LD: history loads through the
// read-only path (0) or L1-bypassing (1)
template <int VEC, int NT, int LD>
__global__ void __launch_bounds__(NT, 2)
k_update(float* __restrict__ x_old, const float* __restrict__ g_old, const float* __restrict__ xn,
         const float* __restrict__ gn, float* __restrict__ Ut, float* __restrict__ Vt,
         float* __restrict__ low_x, float* __restrict__ low_g, const impflow_broyden_state* st, long long d,
         int T, int SL, int expect_nstep) {
  constexpr int kThreads = NT;
  constexpr int kWarps = NT / 32;
  const int do_update = st->do_update;
  const int new_low = st->new_lowest;
  if (!do_update && !new_low) return;  // uniform over the whole grid
  if (expect_nstep >= 0 && st->nstep != expect_nstep) return;   // this iteration's decision kernel was gated off

  cg::cluster_group cluster = cg::this_cluster();
  const int C = cluster.num_blocks();
  const int r = cluster.block_rank();
  const int b = blockIdx.x / C;
  const int k = st->nstep - 1;  // history entries already stored = slot to write (:174)

  extern __shared__ __align__(16) float smem[];
  float* sdx = smem;
  float* sdg = sdx + SL;
  float* sgn = sdg + SL;
  float* wsum = sgn + SL;                 // [3*T][kWarps]
  float* part0 = wsum + 3 * T * kWarps;   // [3*T]
  float* part1 = part0 + 3 * T;           // [2]
  float* tot = part1 + 2;                 // [3*T + 2]

  const long long base = (long long)b * d + (long long)r * SL;
  long long rem = d - (long long)r * SL;
  const int len = rem <= 0 ? 0 : (rem < SL ? (int)rem : SL);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- phase 0: deltas into smem, best-iterate copy (:159-162, :94-99) ----
  if (VEC == 4) {
    constexpr int UN = NT >= 512 ? 2 : 4;       // 16 (8) independent loads per thread in flight
    for (int i0 = tid * 4; i0 < len; i0 += kThreads * 4 * UN) {
      float4 xo[UN], xv[UN], go[UN], gv[UN];
#pragma unroll
      for (int q = 0; q < UN; ++q) {
        const int i = i0 + q * kThreads * 4;
        if (i < len) {
          xo[q] = *reinterpret_cast<const float4*>(x_old + base + i);
          xv[q] = *reinterpret_cast<const float4*>(xn + base + i);
          go[q] = *reinterpret_cast<const float4*>(g_old + base + i);
          gv[q] = *reinterpret_cast<const float4*>(gn + base + i);
        }
      }
#pragma unroll
      for (int q = 0; q < UN; ++q) {
        const int i = i0 + q * kThreads * 4;
        if (i < len) {
          *reinterpret_cast<float4*>(sdx + i) =
              make_float4(xv[q].x - xo[q].x, xv[q].y - xo[q].y, xv[q].z - xo[q].z, xv[q].w - xo[q].w);
          *reinterpret_cast<float4*>(sdg + i) =
              make_float4(gv[q].x - go[q].x, gv[q].y - go[q].y, gv[q].z - go[q].z, gv[q].w - go[q].w);
          *reinterpret_cast<float4*>(sgn + i) = gv[q];
          if (new_low) {
            *reinterpret_cast<float4*>(low_x + base + i) = xv[q];
            *reinterpret_cast<float4*>(low_g + base + i) = gv[q];
          }
        }
      }
    }
  }
  for (int i = tid * VEC; VEC != 4 && i < len; i += kThreads * VEC) {
    if (VEC == 4) {
    } else {
      const float xv = xn[base + i], gv = gn[base + i];
      sdx[i] = xv - x_old[base + i];
      sdg[i] = gv - g_old[base + i];
      sgn[i] = gv;
      if (new_low) {
        low_x[base + i] = xv;
        low_g[base + i] = gv;
      }
    }
  }
  if (!do_update) return;  // uniform
  __syncthreads();

  const float* Ub = Ut + (long long)b * T * d + (long long)r * SL;
  const float* Vb = Vt + (long long)b * T * d + (long long)r * SL;

  // ---- phase 1: a_j = dx.U_j, b_j = V_j.dg, c_j = V_j.gn for j < k (:108,:119) ----
  // Rows in DESCENDING order: phase 2 must accumulate in ascending order and then starts with the rows this pass
  // read last, which are still in L2 when the history in flight exceeds it (same-order passes over a working set
  // larger than the cache hit nothing).  The dots of different rows are independent: results unchanged.
  for (int j = k - 1; j >= 0; --j) {
    const float* uj = Ub + (long long)j * d;
    const float* vj = Vb + (long long)j * d;
    float a = 0.f, bb = 0.f, c = 0.f;
    if (VEC == 4) {
      // the kernel is bound by load latency, not by DRAM bytes: 16 independent 16-byte loads per thread in flight
      // (same per-thread summation order as a plain loop)
      constexpr int UN = NT >= 512 ? 4 : 8;
      for (int i0 = tid * 4; i0 < len; i0 += kThreads * 4 * UN) {
        float4 u4[UN], v4[UN];
#pragma unroll
        for (int q = 0; q < UN; ++q) {
          const int i = i0 + q * kThreads * 4;
          if (i < len) {
            u4[q] = ld_hist<LD>(reinterpret_cast<const float4*>(uj + i));
            v4[q] = ld_hist<LD>(reinterpret_cast<const float4*>(vj + i));
          }
        }
#pragma unroll
        for (int q = 0; q < UN; ++q) {
          const int i = i0 + q * kThreads * 4;
          if (i < len) {
            const float4 dx = *reinterpret_cast<const float4*>(sdx + i);
            const float4 dg = *reinterpret_cast<const float4*>(sdg + i);
            const float4 gg = *reinterpret_cast<const float4*>(sgn + i);
            a += dx.x * u4[q].x + dx.y * u4[q].y + dx.z * u4[q].z + dx.w * u4[q].w;
            bb += v4[q].x * dg.x + v4[q].y * dg.y + v4[q].z * dg.z + v4[q].w * dg.w;
            c += v4[q].x * gg.x + v4[q].y * gg.y + v4[q].z * gg.z + v4[q].w * gg.w;
          }
        }
      }
    } else {
      for (int i = tid; i < len; i += kThreads) {
        const float u1 = uj[i], v1 = vj[i];
        a += sdx[i] * u1;
        bb += v1 * sdg[i];
        c += v1 * sgn[i];
      }
    }
    a = warp_sum(a);
    bb = warp_sum(bb);
    c = warp_sum(c);
    if (lane == 0) {
      wsum[(3 * j + 0) * kWarps + warp] = a;
      wsum[(3 * j + 1) * kWarps + warp] = bb;
      wsum[(3 * j + 2) * kWarps + warp] = c;
    }
  }
  __syncthreads();
  for (int t = tid; t < 3 * k; t += kThreads) {
    float s = 0.f;
    for (int w = 0; w < kWarps; ++w) s += wsum[t * kWarps + w];
    part0[t] = s;
  }
  cluster.sync();
  for (int t = tid; t < 3 * k; t += kThreads) {
    float s = 0.f;
    for (int q = 0; q < C; ++q) s += cluster.map_shared_rank(part0, q)[t];
    tot[t] = s;
  }
  __syncthreads();

  // ---- phase 2: vT, w = matvec(dg), S = sum_j c_j U_j; den = vT.dg, c_k = vT_scrubbed.gn ----
  float den = 0.f, ck = 0.f;
  float* vk = Vt + (long long)b * T * d + (long long)k * d + (long long)r * SL;
  for (int i = tid * VEC; i < len; i += kThreads * VEC) {
    float dx[VEC], dg[VEC], gg[VEC], vT[VEC], w[VEC], S[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      dx[e] = sdx[i + e];
      dg[e] = sdg[i + e];
      gg[e] = sgn[i + e];
      vT[e] = -dx[e];
      w[e] = -dg[e];
      S[e] = 0.f;
    }
    int j = 0;
    if (VEC == 4) {
      constexpr int UJ = NT >= 512 ? 4 : 8;       // 16 (8) independent loads in flight per thread; accumulation order over j unchanged
      for (; j + UJ <= k; j += UJ) {
        float4 u4[UJ], v4[UJ];
#pragma unroll
        for (int q = 0; q < UJ; ++q) {
          u4[q] = ld_hist<LD>(reinterpret_cast<const float4*>(Ub + (long long)(j + q) * d + i));
          v4[q] = ld_hist<LD>(reinterpret_cast<const float4*>(Vb + (long long)(j + q) * d + i));
        }
#pragma unroll
        for (int q = 0; q < UJ; ++q) {
          const float aj = tot[3 * (j + q) + 0], bj = tot[3 * (j + q) + 1], cj = tot[3 * (j + q) + 2];
          const float u1[4] = {u4[q].x, u4[q].y, u4[q].z, u4[q].w};
          const float v1[4] = {v4[q].x, v4[q].y, v4[q].z, v4[q].w};
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            vT[e] += aj * v1[e % 4];
            w[e] += bj * u1[e % 4];
            S[e] += cj * u1[e % 4];
          }
        }
      }
    }
    for (; j < k; ++j) {
      float u1[VEC], v1[VEC];
      if (VEC == 4) {
        const float4 u4 = ld_hist<LD>(reinterpret_cast<const float4*>(Ub + (long long)j * d + i));
        const float4 v4 = ld_hist<LD>(reinterpret_cast<const float4*>(Vb + (long long)j * d + i));
        u1[0] = u4.x; u1[1 % VEC] = u4.y; u1[2 % VEC] = u4.z; u1[3 % VEC] = u4.w;
        v1[0] = v4.x; v1[1 % VEC] = v4.y; v1[2 % VEC] = v4.z; v1[3 % VEC] = v4.w;
      } else {
        u1[0] = __ldg(Ub + (long long)j * d + i);
        v1[0] = __ldg(Vb + (long long)j * d + i);
      }
      const float aj = tot[3 * j + 0], bj = tot[3 * j + 1], cj = tot[3 * j + 2];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        vT[e] += aj * v1[e];
        w[e] += bj * u1[e];
        S[e] += cj * u1[e];
      }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      den += vT[e] * dg[e];                          // unscrubbed vT in the denominator (:176)
      vT[e] = (vT[e] != vT[e]) ? 0.f : vT[e];        // :177
      ck += vT[e] * gg[e];
      sdx[i + e] = dx[e] - w[e];                     // numerator of u
      sdg[i + e] = S[e];
    }
    if (VEC == 4) {
      *reinterpret_cast<float4*>(vk + i) = make_float4(vT[0], vT[1 % VEC], vT[2 % VEC], vT[3 % VEC]);  // :179
    } else {
      vk[i] = vT[0];
    }
  }
  den = warp_sum(den);
  ck = warp_sum(ck);
  if (lane == 0) {
    wsum[warp] = den;
    wsum[kWarps + warp] = ck;
  }
  __syncthreads();
  if (tid < 2) {
    float s = 0.f;
    for (int w = 0; w < kWarps; ++w) s += wsum[tid * kWarps + w];
    part1[tid] = s;
  }
  cluster.sync();
  if (tid < 2) {
    float s = 0.f;
    for (int q = 0; q < C; ++q) s += cluster.map_shared_rank(part1, q)[tid];
    tot[3 * T + tid] = s;
  }
  __syncthreads();
  const float den_t = tot[3 * T + 0], ck_t = tot[3 * T + 1];

  // ---- phase 3: u (scrub), next direction and next iterate (:176-181, :90) ----
  float* uk = Ut + (long long)b * T * d + (long long)k * d + (long long)r * SL;
  constexpr int U3 = 8;       // VEC == 4: the iterate is fetched 8 groups ahead (one round of load latency per slice)
  float4 xpre[U3];
  for (int i = tid * VEC, it = 0; i < len; i += kThreads * VEC, ++it) {
    float u[VEC], xnext[VEC], xv[VEC];
    if (VEC == 4) {
      if ((it % U3) == 0) {
#pragma unroll
        for (int q = 0; q < U3; ++q) {
          const int iq = i + q * kThreads * 4;
          if (iq < len) xpre[q] = *reinterpret_cast<const float4*>(xn + base + iq);
        }
      }
      float4 x4 = xpre[0];
#pragma unroll
      for (int q = 1; q < U3; ++q) x4 = ((it % U3) == q) ? xpre[q] : x4;
      xv[0] = x4.x; xv[1 % VEC] = x4.y; xv[2 % VEC] = x4.z; xv[3 % VEC] = x4.w;
    } else {
      xv[0] = xn[base + i];
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float q = sdx[i + e] / den_t;
      q = (q != q) ? 0.f : q;                        // :178
      u[e] = q;
      const float upd = -((-sgn[i + e]) + (sdg[i + e] + q * ck_t));  // -matvec(U[:nstep], V[:nstep], gx) :181
      xnext[e] = xv[e] + upd;
    }
    if (VEC == 4) {
      *reinterpret_cast<float4*>(uk + i) = make_float4(u[0], u[1 % VEC], u[2 % VEC], u[3 % VEC]);   // :180
      *reinterpret_cast<float4*>(x_old + base + i) =
          make_float4(xnext[0], xnext[1 % VEC], xnext[2 % VEC], xnext[3 % VEC]);
    } else {
      uk[i] = u[0];
      x_old[base + i] = xnext[0];
    }
  }
  cluster.sync();  // peers may still be reading part1 through DSMEM
}

// ------------------------------------------------------------------------------------------
// rank-1 update, cluster per sample, history read from DRAM ONCE (d % 4 == 0).
// The coefficient of history row j in phase 2 (a_j, b_j, c_j) depends on row j alone, so the history is
// walked in chunks of G rows: dots of the chunk (first read: DRAM) -> cluster exchange -> accumulation of
// the chunk into vT, w, S (second read: the rows a CTA touched a few microseconds ago, still in L2 as long
// as `resident CTAs x 2 G SL 4` bytes fit).  k_update<4> does the same arithmetic with G = k; per-row
// reduction order, order of the accumulation over j and the expressions are unchanged, so the results are
// bit-identical.  The running sums live in registers (IT float4 groups per thread = SL / 1024).
// ------------------------------------------------------------------------------------------
template <int IT>
__global__ void __launch_bounds__(kThreads)
k_update_chunked(float* __restrict__ x_old, const float* __restrict__ g_old, const float* __restrict__ xn,
                 const float* __restrict__ gn, float* __restrict__ Ut, float* __restrict__ Vt,
                 float* __restrict__ low_x, float* __restrict__ low_g, const impflow_broyden_state* st, long long d,
                 int T, int SL, int expect_nstep, int G) {
  const int do_update = st->do_update;
  const int new_low = st->new_lowest;
  if (!do_update && !new_low) return;  // uniform over the whole grid
  if (expect_nstep >= 0 && st->nstep != expect_nstep) return;

  cg::cluster_group cluster = cg::this_cluster();
  const int C = cluster.num_blocks();
  const int r = cluster.block_rank();
  const int b = blockIdx.x / C;
  const int k = st->nstep - 1;

  extern __shared__ __align__(16) float smem[];
  float* sdx = smem;
  float* sdg = sdx + SL;
  float* sgn = sdg + SL;
  float* wsum = sgn + SL;                 // [3*T][kWarps]
  float* part0 = wsum + 3 * T * kWarps;   // [3*T]
  float* part1 = part0 + 3 * T;           // [2]
  float* tot = part1 + 2;                 // [3*T + 2]

  const long long base = (long long)b * d + (long long)r * SL;
  long long rem = d - (long long)r * SL;
  const int len = rem <= 0 ? 0 : (rem < SL ? (int)rem : SL);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- phase 0 ----
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const int i = tid * 4 + it * kThreads * 4;
    if (i < len) {
      const float4 xo = *reinterpret_cast<const float4*>(x_old + base + i);
      const float4 xv = *reinterpret_cast<const float4*>(xn + base + i);
      const float4 go = *reinterpret_cast<const float4*>(g_old + base + i);
      const float4 gv = *reinterpret_cast<const float4*>(gn + base + i);
      *reinterpret_cast<float4*>(sdx + i) = make_float4(xv.x - xo.x, xv.y - xo.y, xv.z - xo.z, xv.w - xo.w);
      *reinterpret_cast<float4*>(sdg + i) = make_float4(gv.x - go.x, gv.y - go.y, gv.z - go.z, gv.w - go.w);
      *reinterpret_cast<float4*>(sgn + i) = gv;
      if (new_low) {
        *reinterpret_cast<float4*>(low_x + base + i) = xv;
        *reinterpret_cast<float4*>(low_g + base + i) = gv;
      }
    }
  }
  if (!do_update) return;  // uniform
  __syncthreads();

  const float* Ub = Ut + (long long)b * T * d + (long long)r * SL;
  const float* Vb = Vt + (long long)b * T * d + (long long)r * SL;

  float vT[IT][4], w[IT][4], S[IT][4];
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const int i = tid * 4 + it * kThreads * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      vT[it][e] = (i < len) ? -sdx[i + e] : 0.f;
      w[it][e] = (i < len) ? -sdg[i + e] : 0.f;
      S[it][e] = 0.f;
    }
  }

  for (int j0 = 0; j0 < k; j0 += G) {
    const int j1 = (j0 + G < k) ? j0 + G : k;
    // ---- phase 1 of the chunk: a_j = dx.U_j, b_j = V_j.dg, c_j = V_j.gn (:108,:119) ----
    for (int j = j0; j < j1; ++j) {
      const float* uj = Ub + (long long)j * d;
      const float* vj = Vb + (long long)j * d;
      float4 u4[IT], v4[IT];
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int i = tid * 4 + it * kThreads * 4;
        if (i < len) {
          u4[it] = __ldg(reinterpret_cast<const float4*>(uj + i));
          v4[it] = __ldg(reinterpret_cast<const float4*>(vj + i));
        }
      }
      float a = 0.f, bb = 0.f, c = 0.f;
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int i = tid * 4 + it * kThreads * 4;
        if (i < len) {
          const float4 dx = *reinterpret_cast<const float4*>(sdx + i);
          const float4 dg = *reinterpret_cast<const float4*>(sdg + i);
          const float4 gg = *reinterpret_cast<const float4*>(sgn + i);
          a += dx.x * u4[it].x + dx.y * u4[it].y + dx.z * u4[it].z + dx.w * u4[it].w;
          bb += v4[it].x * dg.x + v4[it].y * dg.y + v4[it].z * dg.z + v4[it].w * dg.w;
          c += v4[it].x * gg.x + v4[it].y * gg.y + v4[it].z * gg.z + v4[it].w * gg.w;
        }
      }
      a = warp_sum(a);
      bb = warp_sum(bb);
      c = warp_sum(c);
      if (lane == 0) {
        wsum[(3 * j + 0) * kWarps + warp] = a;
        wsum[(3 * j + 1) * kWarps + warp] = bb;
        wsum[(3 * j + 2) * kWarps + warp] = c;
      }
    }
    __syncthreads();
    for (int t = 3 * j0 + tid; t < 3 * j1; t += kThreads) {
      float s = 0.f;
      for (int q = 0; q < kWarps; ++q) s += wsum[t * kWarps + q];
      part0[t] = s;
    }
    cluster.sync();      // slots 3*j0 .. 3*j1 are written once per launch: no second barrier per chunk
    for (int t = 3 * j0 + tid; t < 3 * j1; t += kThreads) {
      float s = 0.f;
      for (int q = 0; q < C; ++q) s += cluster.map_shared_rank(part0, q)[t];
      tot[t] = s;
    }
    __syncthreads();
    // ---- phase 2 of the chunk: vT += a_j V_j, w += b_j U_j, S += c_j U_j (rows re-read from L2) ----
    for (int j = j0; j < j1; ++j) {
      const float* uj = Ub + (long long)j * d;
      const float* vj = Vb + (long long)j * d;
      const float aj = tot[3 * j + 0], bj = tot[3 * j + 1], cj = tot[3 * j + 2];
      float4 u4[IT], v4[IT];
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int i = tid * 4 + it * kThreads * 4;
        if (i < len) {
          u4[it] = __ldg(reinterpret_cast<const float4*>(uj + i));
          v4[it] = __ldg(reinterpret_cast<const float4*>(vj + i));
        }
      }
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int i = tid * 4 + it * kThreads * 4;
        if (i < len) {
          const float u1[4] = {u4[it].x, u4[it].y, u4[it].z, u4[it].w};
          const float v1[4] = {v4[it].x, v4[it].y, v4[it].z, v4[it].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            vT[it][e] += aj * v1[e];
            w[it][e] += bj * u1[e];
            S[it][e] += cj * u1[e];
          }
        }
      }
    }
  }

  // ---- den = vT.dg, c_k = vT_scrubbed.gn; numerator of u and S into smem; store vT (:176-179) ----
  float den = 0.f, ck = 0.f;
  float* vk = Vt + (long long)b * T * d + (long long)k * d + (long long)r * SL;
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const int i = tid * 4 + it * kThreads * 4;
    if (i < len) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float dx = sdx[i + e], dg = sdg[i + e], gg = sgn[i + e];
        den += vT[it][e] * dg;                                         // unscrubbed vT in the denominator (:176)
        vT[it][e] = (vT[it][e] != vT[it][e]) ? 0.f : vT[it][e];        // :177
        ck += vT[it][e] * gg;
        sdx[i + e] = dx - w[it][e];                                    // numerator of u
        sdg[i + e] = S[it][e];
      }
      *reinterpret_cast<float4*>(vk + i) = make_float4(vT[it][0], vT[it][1], vT[it][2], vT[it][3]);  // :179
    }
  }
  den = warp_sum(den);
  ck = warp_sum(ck);
  if (lane == 0) {
    wsum[warp] = den;
    wsum[kWarps + warp] = ck;
  }
  __syncthreads();
  if (tid < 2) {
    float s = 0.f;
    for (int q = 0; q < kWarps; ++q) s += wsum[tid * kWarps + q];
    part1[tid] = s;
  }
  cluster.sync();
  if (tid < 2) {
    float s = 0.f;
    for (int q = 0; q < C; ++q) s += cluster.map_shared_rank(part1, q)[tid];
    tot[3 * T + tid] = s;
  }
  __syncthreads();
  const float den_t = tot[3 * T + 0], ck_t = tot[3 * T + 1];

  // ---- phase 3: u (scrub), next direction and next iterate (:176-181, :90) ----
  float* uk = Ut + (long long)b * T * d + (long long)k * d + (long long)r * SL;
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const int i = tid * 4 + it * kThreads * 4;
    if (i < len) {
      const float4 x4 = *reinterpret_cast<const float4*>(xn + base + i);
      const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
      float u[4], xnext[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float q = sdx[i + e] / den_t;
        q = (q != q) ? 0.f : q;                        // :178
        u[e] = q;
        const float upd = -((-sgn[i + e]) + (sdg[i + e] + q * ck_t));  // -matvec(U[:nstep], V[:nstep], gx) :181
        xnext[e] = xv[e] + upd;
      }
      *reinterpret_cast<float4*>(uk + i) = make_float4(u[0], u[1], u[2], u[3]);   // :180
      *reinterpret_cast<float4*>(x_old + base + i) = make_float4(xnext[0], xnext[1], xnext[2], xnext[3]);
    }
  }
  cluster.sync();  // peers may still be reading part1 through DSMEM
}

// ------------------------------------------------------------------------------------------
// rank-1 update for histories that do not fit in L2 (classifier shapes: B = 128, d = 65536, 2 GB): every history
// row crosses the SM boundary ONCE.  Cluster per sample, slices of <= 4096 floats.  A producer warp streams the
// (U_j, V_j) slices through a ring of kStreamStages shared-memory stages with 1-D bulk copies (cp.async.bulk,
// completion counted on an mbarrier); the 8 compute warps take the three dots of row j from shared memory, PUSH the
// block sums into every peer's shared memory (st.shared::cluster + remote mbarrier arrive: no blocking cluster
// barrier, several rows in flight) and, kStreamAhead rows later, accumulate row j into vT, w, S from the SAME stage.
// Reduction order per row: thread partial (ascending i) -> warp tree -> warps ascending -> ranks ascending, as in
// k_update<4>; only the slice length differs (so results agree to round-off, not bit for bit).
// dynamic smem: sdx[SL] sdg[SL] | ring[kStreamStages][2][SL] | parts[T][3][C] | wsum[3][8] bsum[4] part1[2] tot1[2]
//               | full[kStreamStages] empty[kStreamStages] xbar[T]
// ------------------------------------------------------------------------------------------
constexpr int kStreamStages = 5;
constexpr int kStreamAhead = 2;
constexpr int kStreamIT = 4;            // float4 groups per thread: slices of <= 4096 floats
constexpr int kStreamThreads = kThreads + 32;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sb_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sb_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void sb_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "SB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni SB_DONE;\n\t"
      "bra.uni SB_WAIT;\n\t"
      "SB_DONE:\n\t"
      "}\n" ::"r"(s_u32(bar)),
      "r"(parity)
      : "memory");
}
// wait for arrivals that peers made with release.cluster (their st.shared::cluster stores become visible)
__device__ __forceinline__ void sb_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "SBC_WAIT:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni SBC_DONE;\n\t"
      "bra.uni SBC_WAIT;\n\t"
      "SBC_DONE:\n\t"
      "}\n" ::"r"(s_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(s_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_remote(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void arrive_remote(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(kStreamThreads, 1)
k_update_stream(float* __restrict__ x_old, const float* __restrict__ g_old, const float* __restrict__ xn,
                const float* __restrict__ gn, float* __restrict__ Ut, float* __restrict__ Vt,
                float* __restrict__ low_x, float* __restrict__ low_g, const impflow_broyden_state* st, long long d,
                int T, int SL, int expect_nstep) {
  constexpr int IT = kStreamIT;
  const int do_update = st->do_update;
  const int new_low = st->new_lowest;
  if (!do_update && !new_low) return;  // uniform over the whole grid
  if (expect_nstep >= 0 && st->nstep != expect_nstep) return;

  cg::cluster_group cluster = cg::this_cluster();
  const int C = cluster.num_blocks();
  const int r = cluster.block_rank();
  const int b = blockIdx.x / C;
  const int k = st->nstep - 1;

  extern __shared__ __align__(16) float smem[];
  float* sdx = smem;
  float* sdg = sdx + SL;
  float* ring = sdg + SL;                                   // [stage][U | V][SL]
  float* parts = ring + (size_t)kStreamStages * 2 * SL;     // [T][3][C]
  float* wsum = parts + (size_t)T * 3 * C;                  // [3][kWarps]
  float* bsum = wsum + 3 * kWarps;                          // [4]
  float* part1 = bsum + 4;                                  // [2]
  float* tot1 = part1 + 2;                                  // [2]
  uint64_t* full = reinterpret_cast<uint64_t*>(tot1 + 2 + (((size_t)T * 3 * C) & 1));   // 8-byte aligned
  uint64_t* empty = full + kStreamStages;
  uint64_t* xbar = empty + kStreamStages;                   // [T]

  const long long base = (long long)b * d + (long long)r * SL;
  long long rem = d - (long long)r * SL;
  const int len = rem <= 0 ? 0 : (rem < SL ? (int)rem : SL);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool compute = tid < kThreads;

  // ---- phase 0: deltas into smem (g_new stays in registers), best-iterate copy ----
  float4 gg[IT];
  if (compute) {
    float4 xo[IT], xv[IT], go[IT];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int i = tid * 4 + it * kThreads * 4;
      if (i < len) {
        xo[it] = *reinterpret_cast<const float4*>(x_old + base + i);
        xv[it] = *reinterpret_cast<const float4*>(xn + base + i);
        go[it] = *reinterpret_cast<const float4*>(g_old + base + i);
        gg[it] = *reinterpret_cast<const float4*>(gn + base + i);
      }
    }
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int i = tid * 4 + it * kThreads * 4;
      if (i < len) {
        *reinterpret_cast<float4*>(sdx + i) =
            make_float4(xv[it].x - xo[it].x, xv[it].y - xo[it].y, xv[it].z - xo[it].z, xv[it].w - xo[it].w);
        *reinterpret_cast<float4*>(sdg + i) =
            make_float4(gg[it].x - go[it].x, gg[it].y - go[it].y, gg[it].z - go[it].z, gg[it].w - go[it].w);
        if (new_low) {
          *reinterpret_cast<float4*>(low_x + base + i) = xv[it];
          *reinterpret_cast<float4*>(low_g + base + i) = gg[it];
        }
      } else {
        gg[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  if (!do_update) return;  // uniform
  if (tid == 0) {
    for (int s = 0; s < kStreamStages; ++s) {
      sb_init(full + s, 1);
      sb_init(empty + s, kWarps);
    }
    for (int j = 0; j < T; ++j) sb_init(xbar + j, C);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster.sync();       // every peer's barriers exist before the first remote arrive

  const float* Ub = Ut + (long long)b * T * d + (long long)r * SL;
  const float* Vb = Vt + (long long)b * T * d + (long long)r * SL;
  const uint32_t row_bytes = (uint32_t)len * 4u;

  float vT[IT][4], w[IT][4], S[IT][4];
  if (!compute) {
    // ---- producer warp: one lane streams the history slices through the ring ----
    if (lane == 0) {
      for (int j = 0; j < k; ++j) {
        const int s = j % kStreamStages;
        if (j >= kStreamStages) sb_wait(empty + s, ((j / kStreamStages) - 1) & 1);
        sb_expect_tx(full + s, 2 * row_bytes);
        if (row_bytes) {
          bulk_g2s(ring + (size_t)s * 2 * SL, Ub + (long long)j * d, row_bytes, full + s);
          bulk_g2s(ring + (size_t)s * 2 * SL + SL, Vb + (long long)j * d, row_bytes, full + s);
        }
      }
    }
  } else {
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int i = tid * 4 + it * kThreads * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        vT[it][e] = (i < len) ? -sdx[i + e] : 0.f;
        w[it][e] = (i < len) ? -sdg[i + e] : 0.f;
        S[it][e] = 0.f;
      }
    }
    // dots of row j from its stage; block sums pushed to every peer (a_j = dx.U_j, b_j = V_j.dg, c_j = V_j.gn)
    auto dots_push = [&](int j) {
      const int s = j % kStreamStages;
      sb_wait(full + s, (j / kStreamStages) & 1);
      const float* us = ring + (size_t)s * 2 * SL;
      const float* vs = us + SL;
      float a = 0.f, bb = 0.f, c = 0.f;
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int i = tid * 4 + it * kThreads * 4;
        if (i < len) {
          const float4 u4 = *reinterpret_cast<const float4*>(us + i);
          const float4 v4 = *reinterpret_cast<const float4*>(vs + i);
          const float4 dx = *reinterpret_cast<const float4*>(sdx + i);
          const float4 dg = *reinterpret_cast<const float4*>(sdg + i);
          a += dx.x * u4.x + dx.y * u4.y + dx.z * u4.z + dx.w * u4.w;
          bb += v4.x * dg.x + v4.y * dg.y + v4.z * dg.z + v4.w * dg.w;
          c += v4.x * gg[it].x + v4.y * gg[it].y + v4.z * gg[it].z + v4.w * gg[it].w;
        }
      }
      a = warp_sum(a);
      bb = warp_sum(bb);
      c = warp_sum(c);
      compute_bar();      // the previous row's wsum / bsum have been consumed
      if (lane == 0) {
        wsum[0 * kWarps + warp] = a;
        wsum[1 * kWarps + warp] = bb;
        wsum[2 * kWarps + warp] = c;
      }
      compute_bar();
      if (warp == 0) {
        if (lane < 3) {
          float t = 0.f;
          for (int q = 0; q < kWarps; ++q) t += wsum[lane * kWarps + q];
          bsum[lane] = t;
        }
        __syncwarp();
        if (lane < C) {
          float* slot = parts + ((size_t)j * 3) * C + r;      // parts[j][t][source rank]
          st_remote(map_to_rank(slot, lane), bsum[0]);
          st_remote(map_to_rank(slot + C, lane), bsum[1]);
          st_remote(map_to_rank(slot + 2 * C, lane), bsum[2]);
          arrive_remote(map_to_rank(xbar + j, lane));
        }
      }
    };
    const int ahead = k < kStreamAhead ? k : kStreamAhead;
    for (int j = 0; j < ahead; ++j) dots_push(j);
    for (int j = 0; j < k; ++j) {
      if (j + kStreamAhead < k) dots_push(j + kStreamAhead);
      sb_wait_cluster(xbar + j, 0);
      float aj = 0.f, bj = 0.f, cj = 0.f;
      const float* pj = parts + ((size_t)j * 3) * C;
      for (int q = 0; q < C; ++q) {
        aj += pj[q];
        bj += pj[C + q];
        cj += pj[2 * C + q];
      }
      const int s = j % kStreamStages;
      const float* us = ring + (size_t)s * 2 * SL;
      const float* vs = us + SL;
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int i = tid * 4 + it * kThreads * 4;
        if (i < len) {
          const float4 u4 = *reinterpret_cast<const float4*>(us + i);
          const float4 v4 = *reinterpret_cast<const float4*>(vs + i);
          const float u1[4] = {u4.x, u4.y, u4.z, u4.w};
          const float v1[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            vT[it][e] += aj * v1[e];
            w[it][e] += bj * u1[e];
            S[it][e] += cj * u1[e];
          }
        }
      }
      __syncwarp();
      if (lane == 0) sb_arrive(empty + s);      // this warp is done with the stage
    }
  }

  // ---- den = vT.dg, c_k = vT_scrubbed.gn; numerator of u and S into smem; store vT (:176-179) ----
  float* vk = Vt + (long long)b * T * d + (long long)k * d + (long long)r * SL;
  if (compute) {
    float den = 0.f, ck = 0.f;
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int i = tid * 4 + it * kThreads * 4;
      if (i < len) {
        const float g4[4] = {gg[it].x, gg[it].y, gg[it].z, gg[it].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float dx = sdx[i + e], dg = sdg[i + e];
          den += vT[it][e] * dg;                                         // unscrubbed vT in the denominator (:176)
          vT[it][e] = (vT[it][e] != vT[it][e]) ? 0.f : vT[it][e];        // :177
          ck += vT[it][e] * g4[e];
          sdx[i + e] = dx - w[it][e];                                    // numerator of u
          sdg[i + e] = S[it][e];
        }
        *reinterpret_cast<float4*>(vk + i) = make_float4(vT[it][0], vT[it][1], vT[it][2], vT[it][3]);  // :179
      }
    }
    den = warp_sum(den);
    ck = warp_sum(ck);
    compute_bar();
    if (lane == 0) {
      wsum[warp] = den;
      wsum[kWarps + warp] = ck;
    }
    compute_bar();
    if (tid < 2) {
      float t = 0.f;
      for (int q = 0; q < kWarps; ++q) t += wsum[tid * kWarps + q];
      part1[tid] = t;
    }
  }
  cluster.sync();
  if (tid < 2) {
    float t = 0.f;
    for (int q = 0; q < C; ++q) t += cluster.map_shared_rank(part1, q)[tid];
    tot1[tid] = t;
  }
  __syncthreads();
  const float den_t = tot1[0], ck_t = tot1[1];

  // ---- phase 3: u (scrub), next direction and next iterate (:176-181, :90) ----
  float* uk = Ut + (long long)b * T * d + (long long)k * d + (long long)r * SL;
  if (compute) {
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int i = tid * 4 + it * kThreads * 4;
      if (i < len) {
        const float4 x4 = *reinterpret_cast<const float4*>(xn + base + i);
        const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
        const float g4[4] = {gg[it].x, gg[it].y, gg[it].z, gg[it].w};
        float u[4], xnext[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float q = sdx[i + e] / den_t;
          q = (q != q) ? 0.f : q;                        // :178
          u[e] = q;
          const float upd = -((-g4[e]) + (sdg[i + e] + q * ck_t));  // -matvec(U[:nstep], V[:nstep], gx) :181
          xnext[e] = xv[e] + upd;
        }
        *reinterpret_cast<float4*>(uk + i) = make_float4(u[0], u[1], u[2], u[3]);   // :180
        *reinterpret_cast<float4*>(x_old + base + i) = make_float4(xnext[0], xnext[1], xnext[2], xnext[3]);
      }
    }
  }
  cluster.sync();  // peers may still be reading part1 through DSMEM
}

// small-d variant: one warp per sample, d <= 128 (toy / tabular shapes).
__global__ void __launch_bounds__(kThreads)
k_update_small(float* __restrict__ x_old, const float* __restrict__ g_old, const float* __restrict__ xn,
               const float* __restrict__ gn, float* __restrict__ Ut, float* __restrict__ Vt,
               float* __restrict__ low_x, float* __restrict__ low_g, const impflow_broyden_state* st, int B,
               int d, int T, int expect_nstep) {
  const int do_update = st->do_update;
  const int new_low = st->new_lowest;
  if (!do_update && !new_low) return;
  if (expect_nstep >= 0 && st->nstep != expect_nstep) return;
  const int b = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const int k = st->nstep - 1;
  const long long base = (long long)b * d;
  float dx[4], dg[4], gg[4], xv[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = lane + 32 * e;
    if (i < d) {
      xv[e] = xn[base + i];
      gg[e] = gn[base + i];
      dx[e] = xv[e] - x_old[base + i];
      dg[e] = gg[e] - g_old[base + i];
      if (new_low) {
        low_x[base + i] = xv[e];
        low_g[base + i] = gg[e];
      }
    } else {
      xv[e] = gg[e] = dx[e] = dg[e] = 0.f;
    }
  }
  if (!do_update) return;
  const float* Ub = Ut + (long long)b * T * d;
  const float* Vb = Vt + (long long)b * T * d;
  float vT[4], w[4], S[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    vT[e] = -dx[e];
    w[e] = -dg[e];
    S[e] = 0.f;
  }
  for (int j = 0; j < k; ++j) {
    float u1[4], v1[4], a = 0.f, bb = 0.f, c = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = lane + 32 * e;
      u1[e] = i < d ? Ub[(long long)j * d + i] : 0.f;
      v1[e] = i < d ? Vb[(long long)j * d + i] : 0.f;
      a += dx[e] * u1[e];
      bb += v1[e] * dg[e];
      c += v1[e] * gg[e];
    }
    a = warp_sum(a);
    bb = warp_sum(bb);
    c = warp_sum(c);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      vT[e] += a * v1[e];
      w[e] += bb * u1[e];
      S[e] += c * u1[e];
    }
  }
  float den = 0.f, ck = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    den += vT[e] * dg[e];
    vT[e] = (vT[e] != vT[e]) ? 0.f : vT[e];
    ck += vT[e] * gg[e];
  }
  den = warp_sum(den);
  ck = warp_sum(ck);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = lane + 32 * e;
    if (i < d) {
      float u = (dx[e] - w[e]) / den;
      u = (u != u) ? 0.f : u;
      Ut[(long long)b * T * d + (long long)k * d + i] = u;
      Vt[(long long)b * T * d + (long long)k * d + i] = vT[e];
      const float upd = -((-gg[e]) + (S[e] + u * ck));
      x_old[base + i] = xv[e] + upd;
    }
  }
}

// update kernel selection (impflow_broyden_set_chunk), all measured at B = 128, d = 65536 (2 GB of history), 15
// iterations, algorithmic bytes / time against 6543 GB/s (scripts/update_bench.py, profiles/r02_update_bench.txt):
//   -1 / 0  k_update<4, 256>   two passes over the history, 16 loads per thread in flight        0.41  (default)
//   -3      k_update<4, 512>   the same with 16 warps per CTA                                     0.42
//   -2      k_update_stream    every row crosses the SM boundary once (bulk-copy ring, pushes)    0.28
//   n > 0   k_update_chunked   two passes in chunks of n rows, second pass from L2                0.31
// ncu (profiles/r02_ncu_k_update.txt): none of them is DRAM-bound (47 % DRAM throughput at rank 8); the two-pass
// kernel is issue / latency bound at 14.5 resident warps per SM (96 KB of shared memory per CTA), the read-once
// variants pay more for their per-row exchange than they save in traffic.
static int g_update_chunk = -1;

static int pick_splits(int B, long long d) {
  int S = 1;
  while ((long long)B * S < 592 && d / (S * 2) >= 2048 && S < 64) S *= 2;
  return S;
}

static void pick_cluster(int B, long long d, int* C_out, int* SL_out) {
  const int kMaxSlice = 8192;
  int C = 1;
  while (C < 8 && (d + C - 1) / C > kMaxSlice) C *= 2;
  while (C < 8 && (long long)B * C < 296 && (d / (C * 2)) >= 512) C *= 2;
  long long sl = (d + C - 1) / C;
  sl = (sl + 3) / 4 * 4;
  *C_out = C;
  *SL_out = (int)sl;
}

}  // namespace impflow

using namespace impflow;

extern "C" int impflow_broyden_set_chunk(int rows) {
  const int was = g_update_chunk;
  g_update_chunk = rows;
  return was;
}

extern "C" size_t impflow_broyden_state_bytes(void) { return sizeof(impflow_broyden_state); }

extern "C" size_t impflow_broyden_workspace_floats(int B, long long d, int threshold) {
  (void)threshold;
  return (size_t)B * 64;
}

static int launch_norm(const float* g, float* partial, float* sample_sq, float* low_sq,
                       impflow_broyden_state* state, int B, long long d, int init, int gated,
                       BroydenProgress* progress, cudaStream_t s);

// history-streaming update kernel: 0 = launched, -1 = error, 1 = the cluster shape cannot be scheduled here
static int launch_update_stream(float* x_old, const float* g_old, const float* xn, const float* gn, float* Ut,
                                float* Vt, float* low_x, float* low_g, impflow_broyden_state* state, int B,
                                long long d, int threshold, int expect_nstep, cudaStream_t s) {
  int C = 1;
  while (C < 16 && (d + C - 1) / C > 1024 * kStreamIT) C *= 2;
  long long sl = (d + C - 1) / C;
  sl = (sl + 3) / 4 * 4;
  const int SL = (int)sl;
  const size_t floats = (size_t)2 * SL + (size_t)kStreamStages * 2 * SL + (size_t)threshold * 3 * C + 3 * kWarps + 4 +
                        2 + 2 + 2;
  const size_t smem = floats * sizeof(float) + sizeof(uint64_t) * (2 * kStreamStages + threshold) + 16;
  static int configured = 0;      // 0 unknown, 1 ok, -1 unsupported
  if (configured == 0) {
    bool ok = cudaFuncSetAttribute(k_update_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) ==
              cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_update_stream, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    configured = ok ? 1 : -1;
    if (!ok) cudaGetLastError();
  }
  if (configured < 0 || smem > 227 * 1024) return 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(B * C));
  cfg.blockDim = dim3(kStreamThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int placed[17] = {0};    // per cluster size: 0 unknown, 1 schedulable, -1 not
  if (placed[C] == 0) {
    int n = 0;
    placed[C] = (cudaOccupancyMaxActiveClusters(&n, k_update_stream, &cfg) == cudaSuccess && n > 0) ? 1 : -1;
    if (placed[C] < 0) cudaGetLastError();
  }
  if (placed[C] < 0) return 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_update_stream, x_old, g_old, xn, gn, Ut, Vt, low_x, low_g,
                                     (const impflow_broyden_state*)state, d, threshold, SL, expect_nstep);
  if (e != cudaSuccess) {
    set_error("broyden_step: stream-kernel cluster launch failed: %s", cudaGetErrorString(e));
    return -1;
  }
  return check_launch("k_update_stream");
}

extern "C" int impflow_broyden_begin(const float* x0, const float* g0, float* xn, float* low_x, float* low_g,
                                     float* sample_sq, float* low_sq, float* partial,
                                     impflow_broyden_state* state, int B, long long d, int threshold,
                                     double eps_scaled, void* stream) {
  return broyden_begin_ex(x0, g0, xn, low_x, low_g, sample_sq, low_sq, partial, state, B, d, threshold, eps_scaled,
                          nullptr, stream);
}

int impflow::broyden_begin_ex(const float* x0, const float* g0, float* xn, float* low_x, float* low_g,
                              float* sample_sq, float* low_sq, float* partial, impflow_broyden_state* state, int B,
                              long long d, int threshold, double eps_scaled, BroydenProgress* progress,
                              void* stream) {
  IMPFLOW_REQUIRE(threshold >= 1 && threshold <= kMaxT, "broyden: threshold %d not in [1,%d]", threshold, kMaxT);
  IMPFLOW_REQUIRE(B >= 1 && d >= 1, "broyden: empty problem B=%d d=%lld", B, d);
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = (long long)B * d;
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_begin<<<blocks, 256, 0, s>>>(x0, g0, xn, low_x, low_g, n, state, threshold, eps_scaled);
  if (check_launch("k_begin")) return -1;
  return launch_norm(g0, partial, sample_sq, low_sq, state, B, d, 1, 0, progress, s);
}

static int launch_norm(const float* g, float* partial, float* sample_sq, float* low_sq,
                       impflow_broyden_state* state, int B, long long d, int init, int gated,
                       BroydenProgress* progress, cudaStream_t s) {
  const int S = pick_splits(B, d);
  dim3 grid(S, B);
  k_norm_decide<<<grid, kThreads, 0, s>>>(g, partial, sample_sq, low_sq, state, B, d, S, init, gated, progress);
  return check_launch("k_norm_decide");
}

extern "C" int impflow_broyden_step(float* x_old, const float* g_old, const float* xn, const float* gn,
                                    float* Ut, float* Vt, float* low_x, float* low_g, float* sample_sq,
                                    float* low_sq, float* partial, impflow_broyden_state* state, int B,
                                    long long d, int threshold, void* stream) {
  return broyden_step_ex(x_old, g_old, xn, gn, Ut, Vt, low_x, low_g, sample_sq, low_sq, partial, state, B, d,
                         threshold, 0, -1, nullptr, stream);
}

int impflow::broyden_step_ex(float* x_old, const float* g_old, const float* xn, const float* gn, float* Ut,
                             float* Vt, float* low_x, float* low_g, float* sample_sq, float* low_sq, float* partial,
                             impflow_broyden_state* state, int B, long long d, int threshold, int gated,
                             int expect_nstep, BroydenProgress* progress, void* stream) {
  IMPFLOW_REQUIRE(threshold >= 1 && threshold <= kMaxT, "broyden: threshold %d not in [1,%d]", threshold, kMaxT);
  cudaStream_t s = (cudaStream_t)stream;
  if (launch_norm(gn, partial, sample_sq, low_sq, state, B, d, 0, gated, progress, s)) return -1;
  if (d <= 128) {
    const int blocks = (B + kWarps - 1) / kWarps;
    k_update_small<<<blocks, kThreads, 0, s>>>(x_old, g_old, xn, gn, Ut, Vt, low_x, low_g, state, B, (int)d,
                                               threshold, expect_nstep);
    return check_launch("k_update_small");
  }
  IMPFLOW_REQUIRE(d <= 8LL * 8192, "broyden: d=%lld exceeds the cluster kernel limit 65536", d);
  if ((d % 4) == 0 && d >= 2048 && g_update_chunk == -2) {
    const int rc = launch_update_stream(x_old, g_old, xn, gn, Ut, Vt, low_x, low_g, state, B, d, threshold,
                                        expect_nstep, s);
    if (rc <= 0) return rc;     // > 0: this device cannot place the cluster; fall through to the two-pass kernel
  }
  int C, SL;
  pick_cluster(B, d, &C, &SL);
  const size_t smem = sizeof(float) * ((size_t)3 * SL + (size_t)3 * threshold * 16 + 3 * threshold + 2 +
                                       3 * threshold + 2 + 8);      // wsum sized for the 16-warp variant
  const bool vec = (d % 4 == 0);
  // k_update_chunked (history walked in chunks of rows whose second pass hits L2) only on request: measured not
  // faster than the two-pass kernel, both are bound by load latency rather than by DRAM bytes
  const int G = (vec && g_update_chunk > 0) ? g_update_chunk : 0;
  // two-pass kernel variants (impflow_broyden_set_chunk codes -3 / -4 / -5 for A/B): 512 threads per CTA double the
  // loads in flight per SM (the kernel is latency bound at 16 warps per SM); L1-bypassing history loads
  int nt = 256;
  void (*kern)(float*, const float*, const float*, const float*, float*, float*, float*, float*,
               const impflow_broyden_state*, long long, int, int, int) = vec ? k_update<4, 256, 0> : k_update<1, 256, 0>;
  size_t smem_pad = 0;
  if (vec && SL >= 2048 && (g_update_chunk == -3 || g_update_chunk == -4 || g_update_chunk == -6)) {
    nt = 512;
    kern = g_update_chunk == -4 ? k_update<4, 512, 1> : k_update<4, 512, 0>;
    if (g_update_chunk == -6) smem_pad = 120 * 1024;      // one CTA per SM: half the history in flight (fits L2)
  } else if (vec && g_update_chunk == -5) {
    kern = k_update<4, 256, 1>;
  }
  const size_t smem_launch = smem > smem_pad ? smem : smem_pad;
  void (*kern_c)(float*, const float*, const float*, const float*, float*, float*, float*, float*,
                 const impflow_broyden_state*, long long, int, int, int, int) = nullptr;
  if (G > 0) {
    kern_c = SL <= 1024 ? k_update_chunked<1> : SL <= 2048 ? k_update_chunked<2> : SL <= 4096 ? k_update_chunked<4>
                                                                                             : k_update_chunked<8>;
  }
  if ((kern_c ? cudaFuncSetAttribute(kern_c, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_launch)
              : cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_launch)) != cudaSuccess) {
    set_error("broyden_step: cannot set %zu bytes of dynamic shared memory", smem);
    return -1;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(B * C));
  cfg.blockDim = dim3(kern_c ? kThreads : nt);
  cfg.dynamicSmemBytes = smem_launch;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = kern_c ? cudaLaunchKernelEx(&cfg, kern_c, x_old, g_old, xn, gn, Ut, Vt, low_x, low_g,
                                              (const impflow_broyden_state*)state, d, threshold, SL, expect_nstep, G)
                         : cudaLaunchKernelEx(&cfg, kern, x_old, g_old, xn, gn, Ut, Vt, low_x, low_g,
                                              (const impflow_broyden_state*)state, d, threshold, SL, expect_nstep);
  if (e != cudaSuccess) {
    set_error("broyden_step: cluster launch failed: %s", cudaGetErrorString(e));
    return -1;
  }
  return check_launch("k_update");
}
