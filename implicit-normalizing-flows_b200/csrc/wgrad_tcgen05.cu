// Weight-gradient contraction on tcgen05 without transposed copies (sm_100a only):
//
//     dW[N1,N2] = G[Mpix,N1]^T * A[Mpix,N2]        K = all pixels / samples, 3xTF32 (hi/lo planes)
//
// Both operands are row-major with the REDUCTION index (pixels) as the slow dimension, i.e. they are
// "MN-major" for the tensor core.  For 32-bit operands the only MN-major shared-memory layout of tcgen05 is
// SWIZZLE_128B_BASE32B (32-byte chunks permuted inside 128-byte rows with a period of 4 rows); TMA writes
// exactly that with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  Boxes are [32 pixels x 32 channels]: 4 pixel rows x
// 128 bytes per 512-byte atom, consecutive 4-pixel groups 512 B apart (stride-byte-offset), 32-channel column
// blocks one box (4096 B) apart (leading-byte-offset); the instruction descriptor marks A and B as MN-major.
// So the same hi/lo planes that feed the next layer's GEMM as a K-major A operand also feed the weight
// gradient, and the transpose passes of the old path are gone.
//
// Persistent CTAs over (128-row tile of N1, BN-column tile of N2, K slice); split-K partials go to a
// workspace and k_wgrad_reduce sums them in a fixed order (deterministic).  Warp roles as in k_gemm_tc3.
#include "tc_common.cuh"

namespace impflow {

constexpr int WG_THREADS = 256;       // warps 0 TMA, 1 MMA, 2 TMEM alloc, 4..7 epilogue
constexpr int WG_BOX_BYTES = 32 * 128;   // one [32 pixels x 32 channels] box

template <int BN>
struct WgCfg {
  static constexpr int kStages = (BN >= 256) ? 2 : (BN >= 128) ? 3 : 4;
  static constexpr int kGBytes = 4 * WG_BOX_BYTES;            // 128 channels of G per plane
  static constexpr int kABytes = (BN / 32) * WG_BOX_BYTES;    // BN channels of A per plane
  static constexpr int kStageBytes = 2 * kGBytes + 2 * kABytes;
  static constexpr int kTmemCols = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

// MN-major, SWIZZLE_128B_BASE32B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, layout type 1):
// LBO = distance between 32-channel column blocks, SBO = distance between 4-row (K) groups
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1)
k_wgrad_tc3(const __grid_constant__ CUtensorMap mapGhi, const __grid_constant__ CUtensorMap mapGlo,
            const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo, long long Mpix,
            int N1, int N2, int splits, float* __restrict__ ws, int slice_major) {
  using Cfg = WgCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tfull = bars + 2 * Cfg::kStages;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (int)(Mpix / 32);
  const int m_tiles = (N1 + 127) / 128;
  const int n_tiles = (N2 + BN - 1) / BN;
  // slice-major order: the CTAs that run side by side work on ALL output tiles of the same few K slices, so every
  // operand slice is fetched from DRAM once and shared through L2 (tile-major order read the A operand once per
  // wave: 829 MB per 512x512 launch against 538 MB of operands)
  const int mn_tiles = m_tiles * n_tiles;
  const int num_items = mn_tiles * splits;
  const int kb_per = (num_kb + splits - 1) / splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapGhi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapGlo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapAhi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapAlo)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int mn = slice_major ? item % mn_tiles : item / splits, ks = slice_major ? item / mn_tiles : item % splits;
        const int n1_0 = (mn / n_tiles) * 128, n2_0 = (mn % n_tiles) * BN;
        const int kb_end = min(num_kb, (ks + 1) * kb_per);
        for (int kb = ks * kb_per; kb < kb_end; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * Cfg::kStageBytes;
          mbar_expect_tx(&full[stage], Cfg::kStageBytes);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            tma_load_2d(&mapGhi, &full[stage], st + j * WG_BOX_BYTES, n1_0 + 32 * j, kb * 32);
            tma_load_2d(&mapGlo, &full[stage], st + Cfg::kGBytes + j * WG_BOX_BYTES, n1_0 + 32 * j, kb * 32);
          }
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) {
            tma_load_2d(&mapAhi, &full[stage], st + 2 * Cfg::kGBytes + j * WG_BOX_BYTES, n2_0 + 32 * j, kb * 32);
            tma_load_2d(&mapAlo, &full[stage], st + 2 * Cfg::kGBytes + Cfg::kABytes + j * WG_BOX_BYTES, n2_0 + 32 * j,
                        kb * 32);
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // c = F32, a = b = TF32, both operands MN-major (bits 15, 16), N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int ks = slice_major ? item / mn_tiles : item % splits;
      const int kb_begin = ks * kb_per, kb_end = min(num_kb, kb_begin + kb_per);
      mbar_wait(tempty, acc_phase ^ 1);
      tc_fence_after();
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t g_hi = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t g_lo = g_hi + Cfg::kGBytes;
          const uint32_t a_hi = g_hi + 2 * Cfg::kGBytes;
          const uint32_t a_lo = a_hi + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {            // 8 pixels per MMA: two 512-byte atoms along K
            const uint32_t koff = k * 1024;
            const uint64_t dgh = make_mnmajor_sw128_desc(g_hi + koff, WG_BOX_BYTES);
            const uint64_t dgl = make_mnmajor_sw128_desc(g_lo + koff, WG_BOX_BYTES);
            const uint64_t dah = make_mnmajor_sw128_desc(a_hi + koff, WG_BOX_BYTES);
            const uint64_t dal = make_mnmajor_sw128_desc(a_lo + koff, WG_BOX_BYTES);
            umma_tf32(tmem_base, dgl, dah, idesc, (kb != kb_begin || k != 0) ? 1u : 0u);
            umma_tf32(tmem_base, dgh, dal, idesc, 1u);
            umma_tf32(tmem_base, dgh, dah, idesc, 1u);
          }
        }
        __syncwarp();
        if (elect_one()) umma_commit(&empty[stage]);
        __syncwarp();
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(tfull);
      __syncwarp();
      acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int mn = slice_major ? item % mn_tiles : item / splits, ks = slice_major ? item / mn_tiles : item % splits;
      const int n1 = (mn / n_tiles) * 128 + q * 32 + lane;
      const int n2_0 = (mn % n_tiles) * BN;
      mbar_wait(tfull, acc_phase);
      tc_fence_after();
      float* dst_row = ws + ((long long)ks * N1 + n1) * N2;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
        const int n0 = n2_0 + c * 32;
        if (n1 < N1 && n0 < N2) {
          if (n0 + 32 <= N2 && (N2 & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float t8[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) t8[u] = __uint_as_float(r[j + u]);
              st_global_v8(dst_row + n0 + j, t8);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < N2) dst_row[n0 + j] = __uint_as_float(r[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
      acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
  }
}

__global__ void __launch_bounds__(256)
k_wgrad_reduce(const float* __restrict__ ws, float* __restrict__ out, long long total, int splits, long long ldo,
               int N2, int transpose_out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[(long long)s * total + i];     // fixed order: deterministic
    const long long r = i / N2;
    const int c = (int)(i % N2);
    if (transpose_out) {
      out[(long long)c * ldo + r] = acc;
    } else {
      out[r * ldo + c] = acc;
    }
  }
}

static int wg_bn(int N2) { return N2 <= 32 ? 32 : (N2 <= 64 ? 64 : (N2 <= 128 ? 128 : 256)); }

static int wg_splits(long long Mpix, int N1, int N2) {
  const int bn = wg_bn(N2);
  const long long tiles = (long long)((N1 + 127) / 128) * ((N2 + bn - 1) / bn);
  const long long num_kb = Mpix / 32;
  long long s = (296 + tiles - 1) / tiles;
  if (s > num_kb / 8) s = num_kb / 8;
  if (s > 128) s = 128;
  return s < 1 ? 1 : (int)s;
}

static int g_wgrad_slice_major = 1;

template <int BN>
static int launch_wgrad(const CUtensorMap* maps, long long Mpix, int N1, int N2, int splits, float* ws,
                        cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_wgrad_tc3<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<BN>::kSmemBytes) !=
        cudaSuccess) {
      set_error("wgrad_tc: cannot set %d bytes of dynamic shared memory", WgCfg<BN>::kSmemBytes);
      return -1;
    }
    attr_set = true;
  }
  const long long items = (long long)((N1 + 127) / 128) * ((N2 + BN - 1) / BN) * splits;
  const int grid = (int)(items < 148 ? items : 148);
  k_wgrad_tc3<BN><<<grid, WG_THREADS, WgCfg<BN>::kSmemBytes, s>>>(maps[0], maps[1], maps[2], maps[3], Mpix, N1, N2,
                                                                   splits, ws, g_wgrad_slice_major);
  return check_launch("k_wgrad_tc3");
}

}  // namespace impflow

using namespace impflow;

extern "C" int impflow_wgrad_set_slice_major(int on) {
  const int prev = g_wgrad_slice_major;
  g_wgrad_slice_major = on ? 1 : 0;
  return prev;
}

extern "C" size_t impflow_wgrad_tc_workspace_floats(long long Mpix, int N1, int N2) {
  return (size_t)wg_splits(Mpix, N1, N2) * (size_t)N1 * (size_t)N2;
}

extern "C" int impflow_wgrad_tc(const float* G_hi, const float* G_lo, long long ldg, const float* A_hi,
                                const float* A_lo, long long lda, float* out, long long ldo, int transpose_out,
                                long long Mpix, int N1, int N2, float* ws, void* stream) {
  IMPFLOW_REQUIRE(Mpix >= 32 && N1 >= 1 && N2 >= 1, "wgrad_tc: empty problem Mpix=%lld N1=%d N2=%d", Mpix, N1, N2);
  if ((Mpix % 32) != 0 || (ldg % 4) != 0 || (lda % 4) != 0 || ldg < N1 || lda < N2) {
    set_error("wgrad_tc: needs Mpix %% 32 == 0 and 16-byte aligned rows (Mpix=%lld ldg=%lld lda=%lld)", Mpix, ldg,
              lda);
    return -2;
  }
  const uintptr_t al = reinterpret_cast<uintptr_t>(G_hi) | reinterpret_cast<uintptr_t>(G_lo) |
                       reinterpret_cast<uintptr_t>(A_hi) | reinterpret_cast<uintptr_t>(A_lo);
  if (al & 15) {
    set_error("wgrad_tc: operand base pointers must be 16-byte aligned");
    return -2;
  }
  IMPFLOW_REQUIRE(ws != nullptr, "wgrad_tc: workspace missing");
  // maps over [Mpix rows x N channels] planes: inner dimension = channels, boxes of 32 channels x 32 pixels
  CUtensorMap maps[4];
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  if (make_map(&maps[0], G_hi, Mpix, N1, ldg, 32, sw) || make_map(&maps[1], G_lo, Mpix, N1, ldg, 32, sw) ||
      make_map(&maps[2], A_hi, Mpix, N2, lda, 32, sw) || make_map(&maps[3], A_lo, Mpix, N2, lda, 32, sw))
    return -1;
  const int splits = wg_splits(Mpix, N1, N2);
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  switch (wg_bn(N2)) {
    case 32: rc = launch_wgrad<32>(maps, Mpix, N1, N2, splits, ws, s); break;
    case 64: rc = launch_wgrad<64>(maps, Mpix, N1, N2, splits, ws, s); break;
    case 128: rc = launch_wgrad<128>(maps, Mpix, N1, N2, splits, ws, s); break;
    default: rc = launch_wgrad<256>(maps, Mpix, N1, N2, splits, ws, s); break;
  }
  if (rc) return rc;
  const long long total = (long long)N1 * N2;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_wgrad_reduce<<<(int)blocks, 256, 0, s>>>(ws, out, total, splits, ldo, N2, transpose_out);
  return check_launch("k_wgrad_reduce");
}
