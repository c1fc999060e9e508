// fp32 CUDA-core GEMM  C[M,N] = A[M,K] * B[N,K]^T  with the fused branch epilogue.
// Exact-fp32 backend of impflow_gemm_nt: used for the small MLP shapes (toy / tabular, K or N
// down to 2) and as the on-device cross-check of the tcgen05 3xTF32 backend.
#include "common.cuh"

namespace impflow {

constexpr int BK = 16;

template <int BM, int BN>
__global__ void __launch_bounds__(256)
k_gemm_nt(const float* __restrict__ A, long long lda, const float* __restrict__ Bm, long long ldb, long long M,
          int N, int K, Epilogue ep_in, int vecA, int vecB) {
  const Epilogue ep = resolve_beta(ep_in);
  constexpr int TM = BM / 16, TN = BN / 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---- load tiles (K-major rows -> transposed smem) ----
#pragma unroll
    for (int it = 0; it < BM / 64; ++it) {
      const int r = (tid >> 2) + it * 64, q = tid & 3;
      const long long gm = m0 + r;
      const int gk = k0 + q * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gm < M) {
        if (vecA && gk + 3 < K) {
          const float4 t = *reinterpret_cast<const float4*>(A + gm * lda + gk);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (gk + e < K) v[e] = A[gm * lda + gk + e];
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) As[q * 4 + e][r] = v[e];
    }
#pragma unroll
    for (int it = 0; it < BN / 64; ++it) {
      const int r = (tid >> 2) + it * 64, q = tid & 3;
      const int gn = n0 + r;
      const int gk = k0 + q * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gn < N) {
        if (vecB && gk + 3 < K) {
          const float4 t = *reinterpret_cast<const float4*>(Bm + (long long)gn * ldb + gk);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (gk + e < K) v[e] = Bm[(long long)gn * ldb + gk + e];
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) Bs[q * 4 + e][r] = v[e];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&As[kk][(i / 4) * 64 + ty * 4]);
        a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&Bs[kk][(j / 4) * 64 + tx * 4]);
        b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long m = m0 + (i / 4) * 64 + ty * 4 + (i % 4);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (j / 4) * 64 + tx * 4 + (j % 4);
      if (n < N) epilogue_store(ep, m, n, acc[i][j]);
    }
  }
}

// General-stride variant for the autograd primitives of the small MLP / classifier shapes:
//   C[m,n] = sum_k A[m*sAm + k*sAk] * B[n*sBn + k*sBk] (+ bias[n])
// so that A B, A^T B, A B^T and A^T B^T all run without materialising a transpose (the derivative of a GEMM
// is a GEMM with transposed operands; with double backward the transposes used to be 25 % of all launches).
template <int BM, bool EP = false>      // rows per CTA: 64, or 32 / 16 when the problem has few row tiles (more CTAs in flight)
__global__ void __launch_bounds__(256)                      // EP: the fused branch epilogue of k_gemm_nt instead of (+ bias)
k_gemm_strided(const float* __restrict__ A, long long sAm, long long sAk, const float* __restrict__ Bm, long long sBn,
               long long sBk, const float* __restrict__ bias, float* __restrict__ out, long long ldc, long long M,
               int N, int K, Epilogue ep_in = Epilogue()) {
  const Epilogue ep = EP ? resolve_beta(ep_in) : ep_in;
  constexpr int BN = 64, TM = BM / 16, TN = 4;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  // element e of the tile = (row r, k): consecutive threads follow the contiguous dimension of the operand
  const bool a_m_contig = sAm == 1, b_n_contig = sBn == 1;
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int e = tid; e < BM * BK; e += 256) {
      const int r = a_m_contig ? (e % BM) : (e / BK), kk = a_m_contig ? (e / BM) : (e % BK);
      const long long gm = m0 + r;
      const int gk = k0 + kk;
      As[kk][r] = (gm < M && gk < K) ? A[gm * sAm + (long long)gk * sAk] : 0.f;
    }
#pragma unroll
    for (int e = tid; e < BN * BK; e += 256) {
      const int r = b_n_contig ? (e % BN) : (e / BK), kk = b_n_contig ? (e / BN) : (e % BK);
      const int gn = n0 + r;
      const int gk = k0 + kk;
      Bs[kk][r] = (gn < N && gk < K) ? Bm[(long long)gn * sBn + (long long)gk * sBk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
      const float4 tb = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float b[4] = {tb.x, tb.y, tb.z, tb.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      if (EP) epilogue_store(ep, m, n, acc[i][j]);
      else out[m * ldc + n] = acc[i][j] + (bias != nullptr ? bias[n] : 0.f);
    }
  }
}

// Weight-gradient contraction of the small / ragged shapes (MLP flows: N1, N2 <= 128, rows = n x batch, any count):
//     dW[N1,N2] = G[M,N1]^T A[M,N2]
// straight from the row-major operands (no transposed copies) and SPLIT ALONG THE ROWS: one 64x64 output tile would
// otherwise walk all M rows on a single SM (93 us per launch at M = 6000; 858 launches = half of the GPU time of a
// tabular step).  grid = (column tiles, row tiles, splits); the partials are summed in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
k_wgrad_simt(const float* __restrict__ G, long long ldg, const float* __restrict__ A, long long lda,
             float* __restrict__ ws, long long M, int N1, int N2, long long rows_per_split) {
  constexpr int BM = 64, BN = 64, TM = 4, TN = 4;
  __shared__ float Gs[BK][BM + 4];
  __shared__ float As[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int n1_0 = blockIdx.y * BM, n2_0 = blockIdx.x * BN;
  const long long m_begin = (long long)blockIdx.z * rows_per_split;
  const long long m_end = m_begin + rows_per_split < M ? m_begin + rows_per_split : M;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  for (long long m0 = m_begin; m0 < m_end; m0 += BK) {
#pragma unroll
    for (int e = tid; e < BM * BK; e += 256) {       // consecutive threads follow the channel (contiguous) dimension
      const int r = e % BM, kk = e / BM;
      const long long gm = m0 + kk;
      Gs[kk][r] = (gm < m_end && n1_0 + r < N1) ? G[gm * ldg + n1_0 + r] : 0.f;
      As[kk][r] = (gm < m_end && n2_0 + r < N2) ? A[gm * lda + n2_0 + r] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 ta = *reinterpret_cast<const float4*>(&Gs[kk][ty * 4]);
      const float4 tb = *reinterpret_cast<const float4*>(&As[kk][tx * 4]);
      const float a[4] = {ta.x, ta.y, ta.z, ta.w};
      const float b[4] = {tb.x, tb.y, tb.z, tb.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dst = ws + (long long)blockIdx.z * N1 * N2;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int n1 = n1_0 + ty * 4 + i;
    if (n1 >= N1) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n2 = n2_0 + tx * 4 + j;
      if (n2 < N2) dst[(long long)n1 * N2 + n2] = acc[i][j];
    }
  }
}

__global__ void __launch_bounds__(256)
k_wgrad_simt_reduce(const float* __restrict__ ws, float* __restrict__ out, long long total, int splits, long long ldo,
                    int N2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[(long long)s * total + i];     // fixed order
    out[(i / N2) * ldo + (i % N2)] = acc;
  }
}

int wgrad_simt_splits(long long M, int N1, int N2) {
  const long long tiles = (long long)((N1 + 63) / 64) * ((N2 + 63) / 64);
  long long s = (444 + tiles - 1) / tiles;           // about three CTAs per SM
  const long long max_s = (M + 63) / 64;             // at least 64 rows per split
  if (s > max_s) s = max_s;
  if (s > 256) s = 256;
  return s < 1 ? 1 : (int)s;
}

int wgrad_simt(const float* G, long long ldg, const float* A, long long lda, float* out, long long ldo, long long M,
               int N1, int N2, float* ws, cudaStream_t s) {
  const int splits = wgrad_simt_splits(M, N1, N2);
  long long rows = (M + splits - 1) / splits;
  rows = (rows + BK - 1) / BK * BK;
  dim3 grid((unsigned)((N2 + 63) / 64), (unsigned)((N1 + 63) / 64), (unsigned)splits);
  k_wgrad_simt<<<grid, 256, 0, s>>>(G, ldg, A, lda, ws, M, N1, N2, rows);
  if (check_launch("k_wgrad_simt")) return -1;
  const long long total = (long long)N1 * N2;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  k_wgrad_simt_reduce<<<(int)blocks, 256, 0, s>>>(ws, out, total, splits, ldo, N2);
  return check_launch("k_wgrad_simt_reduce");
}

int gemm_strided_simt(const float* A, long long sAm, long long sAk, const float* Bm, long long sBn, long long sBk,
                      const float* bias, float* out, long long ldc, long long M, int N, int K, cudaStream_t s) {
  const long long col_tiles = (N + 63) / 64;
  // enough CTAs to cover the SMs: shrink the row tile while the grid is below ~2 CTAs per SM
  int bm = 64;
  if (((M + 63) / 64) * col_tiles < 296) bm = 32;
  if (((M + 31) / 32) * col_tiles < 296) bm = 16;
  dim3 grid((unsigned)col_tiles, (unsigned)((M + bm - 1) / bm));
  IMPFLOW_REQUIRE(grid.y <= 65535, "gemm_strided: M=%lld too large", M);
  if (bm == 64) k_gemm_strided<64><<<grid, 256, 0, s>>>(A, sAm, sAk, Bm, sBn, sBk, bias, out, ldc, M, N, K);
  else if (bm == 32) k_gemm_strided<32><<<grid, 256, 0, s>>>(A, sAm, sAk, Bm, sBn, sBk, bias, out, ldc, M, N, K);
  else k_gemm_strided<16><<<grid, 256, 0, s>>>(A, sAm, sAk, Bm, sBn, sBk, bias, out, ldc, M, N, K);
  return check_launch("k_gemm_strided");
}

int gemm_nt_simt(const float* A, long long lda, const float* Bm, long long ldb, long long M, int N, int K,
                 const Epilogue& ep, cudaStream_t s) {
  const int vecA = ((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (lda % 4) == 0) ? 1 : 0;
  const int vecB = ((reinterpret_cast<uintptr_t>(Bm) & 15) == 0 && (ldb % 4) == 0) ? 1 : 0;
  const long long tiles_big = ((M + 127) / 128) * ((N + 127) / 128);
  const long long col_tiles = (N + 63) / 64;
  if (((M + 63) / 64) * col_tiles < 148) {
    // few 64x64 tiles (MLP layers of the toy / tabular flows: M = 1000, N = 128 is 32 tiles on 148 SMs, 14 us): the
    // element-strided kernel with 32- or 16-row tiles and the same fused epilogue fills the machine (6 us)
    const int bm = (((M + 31) / 32) * col_tiles < 296) ? 16 : 32;
    dim3 grid((unsigned)col_tiles, (unsigned)((M + bm - 1) / bm));
    IMPFLOW_REQUIRE(grid.y <= 65535, "gemm_nt: M=%lld too large", M);
    if (bm == 32)
      k_gemm_strided<32, true><<<grid, 256, 0, s>>>(A, lda, 1, Bm, ldb, 1, nullptr, nullptr, ep.ldc, M, N, K, ep);
    else
      k_gemm_strided<16, true><<<grid, 256, 0, s>>>(A, lda, 1, Bm, ldb, 1, nullptr, nullptr, ep.ldc, M, N, K, ep);
    return check_launch("k_gemm_strided<ep>");
  }
  if (tiles_big >= 148 && N > 64) {
    dim3 grid((N + 127) / 128, (unsigned)((M + 127) / 128));
    IMPFLOW_REQUIRE(grid.y <= 65535, "gemm_nt: M=%lld too large", M);
    k_gemm_nt<128, 128><<<grid, 256, 0, s>>>(A, lda, Bm, ldb, M, N, K, ep, vecA, vecB);
  } else {
    dim3 grid((N + 63) / 64, (unsigned)((M + 63) / 64));
    IMPFLOW_REQUIRE(grid.y <= 65535, "gemm_nt: M=%lld too large", M);
    k_gemm_nt<64, 64><<<grid, 256, 0, s>>>(A, lda, Bm, ldb, M, N, K, ep, vecA, vecB);
  }
  return check_launch("k_gemm_nt");
}

}  // namespace impflow

using namespace impflow;

extern "C" size_t impflow_wgrad_simt_workspace_floats(long long M, int N1, int N2) {
  return (size_t)wgrad_simt_splits(M, N1, N2) * (size_t)N1 * (size_t)N2;
}

extern "C" int impflow_wgrad_simt(const float* G, long long ldg, const float* A, long long lda, float* out,
                                  long long ldo, long long M, int N1, int N2, float* ws, void* stream) {
  IMPFLOW_REQUIRE(M >= 1 && N1 >= 1 && N2 >= 1, "wgrad_simt: empty problem M=%lld N1=%d N2=%d", M, N1, N2);
  IMPFLOW_REQUIRE(ldg >= N1 && lda >= N2 && ldo >= N2, "wgrad_simt: row strides too small");
  IMPFLOW_REQUIRE(out != nullptr && ws != nullptr, "wgrad_simt: output or workspace missing");
  return wgrad_simt(G, ldg, A, lda, out, ldo, M, N1, N2, ws, (cudaStream_t)stream);
}
