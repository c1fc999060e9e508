// Fused residual-branch tile kernel (sm_100a): the whole narrow -> wide -> wide -> narrow chain of a
// 3x3 / 1x1 / 3x3 conv branch (implicit_flow.py:359-398) — forward or transposed (vjp) — in ONE launch:
//
//     S  = X0 W1^T                  [M,32] x [C,32]^T      X0 = im2col of the narrow side (9c <= 32)
//     A1 = psi1(S)                  bias + activation        (forward)  |  S * D1 (D1 = act'(pre))  (vjp)
//     H  = A1 W2^T                  [M,C] x [C,C]^T
//     A2 = psi2(H)
//     Y += A2 W3^T                  [M,C] x [N3,C]^T        N3 = 9c <= 32 tap columns (col2im follows)
//
// The C-wide intermediates never touch HBM (the unfused path writes and re-reads them as tf32 hi/lo
// planes: ~1.3 GB per evaluation at the CIFAR scale-0 shape).  Everything is fp32-accurate "3xTF32":
// every operand is a hi + lo pair of tf32 values and each product is lo*hi + hi*lo + hi*hi in fp32.
//
// One CTA per SM, persistent over work items (m_tile of 128 rows, n_half of 256 output channels of
// layer 2).  Tensor memory (512 columns) holds every MMA operand that is produced on chip:
//     [  0,256)  ACC   layer-2 accumulator           [256,288) S0   [288,320) S1 (layer-1 chunks; S1
//     [320,384)  A ring slot 0 (hi 32 | lo 32)                        doubles as the layer-3 accumulator)
//     [384,448)  A ring slot 1                       [448,512) X0 hi | lo
// Warps: 0 = TMA producer (weight chunks: W2 256x32 + W1 32x32 per K chunk, then W3 32x32 chunks;
// 3-stage ring of 72 KB), 1 = MMA issuer (tcgen05.mma kind::tf32, A from TMEM, B from smem),
// 2 = TMEM allocator, 4..19 = transform warps (4 per TMEM lane quarter, 8 columns of a chunk each:
// tcgen05.ld -> psi -> tf32 split -> tcgen05.st).
#include "tile_common.cuh"

namespace impflow {

constexpr int BF_NS = 3;                       // weight stages
constexpr int BF_W2_BYTES = 256 * TC_BK * 4;   // 32 KB per plane
constexpr int BF_W1_BYTES = 32 * TC_BK * 4;    // 4 KB per plane
constexpr int BF_STAGE_BYTES = 2 * BF_W2_BYTES + 2 * BF_W1_BYTES;   // 72 KB
constexpr int BF_SMEM_BYTES = BF_NS * BF_STAGE_BYTES + 1024 + 256;
constexpr int BF_THREADS = 128 + 32 * BF_XF_WARPS;
constexpr uint32_t BF_COL_ACC = 0, BF_COL_S = 256, BF_COL_A = 320, BF_COL_X0 = 448;

struct BranchArgs {
  const float* x0;       // [M, ldx] fp32, first 32 columns used
  long long ldx;
  const float* bias1;    // [C] or null
  const float* bias2;    // [C] or null
  const float* mul1;     // [M, C] or null: psi1 = S * mul1 (already act'(pre))
  const float* mul2;     // [M, C] or null
  float* pre1_out;       // [M, C] or null: S + bias1 (forward, kept for the vjp)
  float* pre2_out;       // [M, C] or null
  float* out;            // [M, ldo]; zero-initialised by the caller when C > 256 (two halves are summed)
  long long ldo;
  long long M;
  int C;
  int N3;
  const float* beta1;    // device scalars softplus(beta) (LipSwish) or null
  const float* beta2;
  const int* gate;       // device gate of the sync-free solver loop (common.cuh) or null
};

template <int ACT>
__global__ void __launch_bounds__(BF_THREADS, 1)
k_branch3(const __grid_constant__ CUtensorMap mapW1hi, const __grid_constant__ CUtensorMap mapW1lo,
          const __grid_constant__ CUtensorMap mapW2hi, const __grid_constant__ CUtensorMap mapW2lo,
          const __grid_constant__ CUtensorMap mapW3hi, const __grid_constant__ CUtensorMap mapW3lo,
          const BranchArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BF_NS * BF_STAGE_BYTES);
  uint64_t* b_full = bars;                 // [NS]  TMA -> MMA
  uint64_t* b_empty = b_full + BF_NS;      // [NS]  MMA -> TMA
  uint64_t* s_full = b_empty + BF_NS;      // [2]   MMA -> transform (layer-1 chunk ready)
  uint64_t* s_empty = s_full + 2;          // [2]   transform -> MMA
  uint64_t* a_full = s_empty + 2;          // [2]   transform -> MMA (operand chunk written)
  uint64_t* a_empty = a_full + 2;          // [2]   MMA -> transform
  uint64_t* x0_full = a_empty + 2;         // transform -> MMA
  uint64_t* x0_empty = x0_full + 1;        // MMA -> transform
  uint64_t* acc_full = x0_empty + 1;       // MMA -> transform (layer-2 accumulator complete)
  uint64_t* acc3_full = acc_full + 1;      // MMA -> transform (layer-3 accumulator complete)
  uint64_t* acc3_empty = acc3_full + 1;    // transform -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc3_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NC = args.C / TC_BK;           // K chunks of layer 2 (= 32-channel chunks of layer 1's output)
  const int n_halves = args.C / 256;
  const int NC3 = 256 / TC_BK;             // K chunks of layer 3 per item
  const long long m_tiles = (args.M + TC_BM - 1) / TC_BM;
  long long num_items = m_tiles * n_halves;

  pdl_trigger();     // col2im behind this kernel may be scheduled as its CTAs retire
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapW1hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapW1lo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapW2hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapW2lo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapW3hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapW3lo)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < BF_NS; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], BF_XF_WARPS);
      mbar_init(&a_full[s], BF_XF_WARPS);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(x0_full, BF_XF_WARPS);
    mbar_init(x0_empty, 1);
    mbar_init(acc_full, 1);
    mbar_init(acc3_full, 1);
    mbar_init(acc3_empty, BF_XF_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();        // the set-up above overlapped the predecessor (im2col); no global memory was touched yet
  if (gate_closed(args.gate)) num_items = 0;   // speculative solver iteration after the loop ended (uniform)

  if (warp == 0) {
    // ================= TMA producer: weight chunks =================
    if (lane == 0) {
      uint32_t gw = 0;
      for (long long item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int nh = (int)(item % n_halves);
        for (int kc = 0; kc < NC; ++kc, ++gw) {
          const int st = gw % BF_NS;
          mbar_wait(&b_empty[st], ((gw / BF_NS) & 1) ^ 1);
          uint8_t* sp = smem + st * BF_STAGE_BYTES;
          mbar_expect_tx(&b_full[st], BF_STAGE_BYTES);
          tma_load_2d(&mapW2hi, &b_full[st], sp, kc * TC_BK, nh * 256);
          tma_load_2d(&mapW2lo, &b_full[st], sp + BF_W2_BYTES, kc * TC_BK, nh * 256);
          tma_load_2d(&mapW1hi, &b_full[st], sp + 2 * BF_W2_BYTES, 0, kc * TC_BK);
          tma_load_2d(&mapW1lo, &b_full[st], sp + 2 * BF_W2_BYTES + BF_W1_BYTES, 0, kc * TC_BK);
        }
        for (int c2 = 0; c2 < NC3; ++c2, ++gw) {
          const int st = gw % BF_NS;
          mbar_wait(&b_empty[st], ((gw / BF_NS) & 1) ^ 1);
          uint8_t* sp = smem + st * BF_STAGE_BYTES;
          mbar_expect_tx(&b_full[st], 2 * BF_W1_BYTES);
          tma_load_2d(&mapW3hi, &b_full[st], sp, nh * 256 + c2 * TC_BK, 0);
          tma_load_2d(&mapW3lo, &b_full[st], sp + BF_W1_BYTES, nh * 256 + c2 * TC_BK, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t idesc256 = idesc_base | ((uint32_t)(256 >> 3) << 17);
    const uint32_t idesc32 = idesc_base | ((uint32_t)(32 >> 3) << 17);
    const uint32_t t_acc = tmem_base + BF_COL_ACC;
    const uint32_t t_x0 = tmem_base + BF_COL_X0;
    uint32_t gw = 0, ga = 0, gs = 0, it = 0;
    for (long long item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      mbar_wait(x0_full, it & 1);
      if (it > 0) mbar_wait(acc3_empty, (it - 1) & 1);    // S1 doubles as the layer-3 accumulator
      tc_fence_after();
      // layer 1 of chunk kc -> S[(gs+kc)&1]
      auto issue_l1 = [&](int kc) {
        const uint32_t w = gw + kc, st = w % BF_NS;
        mbar_wait(&b_full[st], (w / BF_NS) & 1);
        const uint32_t sc = gs + kc, buf = sc & 1;
        mbar_wait(&s_empty[buf], ((sc >> 1) & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t w1_hi = smem_u32(smem + st * BF_STAGE_BYTES + 2 * BF_W2_BYTES);
          const uint32_t w1_lo = w1_hi + BF_W1_BYTES;
          const uint32_t t_s = tmem_base + BF_COL_S + buf * 32;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t dbh = make_kmajor_sw128_desc(w1_hi + k * 32);
            const uint64_t dbl = make_kmajor_sw128_desc(w1_lo + k * 32);
            umma_tf32_ts(t_s, t_x0 + 32 + k * 8, dbh, idesc32, k != 0 ? 1u : 0u);   // lo * hi
            umma_tf32_ts(t_s, t_x0 + k * 8, dbl, idesc32, 1u);                        // hi * lo
            umma_tf32_ts(t_s, t_x0 + k * 8, dbh, idesc32, 1u);                        // hi * hi
          }
          umma_commit(&s_full[buf]);
          if (kc == NC - 1) umma_commit(x0_empty);
        }
        __syncwarp();
      };
      issue_l1(0);
      for (int kc = 0; kc < NC; ++kc) {
        if (kc + 1 < NC) issue_l1(kc + 1);
        const uint32_t w = gw + kc, st = w % BF_NS;
        const uint32_t ac = ga + kc, slot = ac & 1;
        mbar_wait(&a_full[slot], (ac >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t w2_hi = smem_u32(smem + st * BF_STAGE_BYTES);
          const uint32_t w2_lo = w2_hi + BF_W2_BYTES;
          const uint32_t t_a = tmem_base + BF_COL_A + slot * 64;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t dbh = make_kmajor_sw128_desc(w2_hi + k * 32);
            const uint64_t dbl = make_kmajor_sw128_desc(w2_lo + k * 32);
            umma_tf32_ts(t_acc, t_a + 32 + k * 8, dbh, idesc256, (kc != 0 || k != 0) ? 1u : 0u);
            umma_tf32_ts(t_acc, t_a + k * 8, dbl, idesc256, 1u);
            umma_tf32_ts(t_acc, t_a + k * 8, dbh, idesc256, 1u);
          }
          umma_commit(&b_empty[st]);
          umma_commit(&a_empty[slot]);
          if (kc == NC - 1) umma_commit(acc_full);
        }
        __syncwarp();
      }
      gw += NC, ga += NC, gs += NC;
      // layer 3: A chunks come from the transformed layer-2 accumulator
      const uint32_t t_acc3 = tmem_base + BF_COL_S + 32;
      for (int c2 = 0; c2 < NC3; ++c2) {
        const uint32_t w = gw + c2, st = w % BF_NS;
        mbar_wait(&b_full[st], (w / BF_NS) & 1);
        const uint32_t ac = ga + c2, slot = ac & 1;
        mbar_wait(&a_full[slot], (ac >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t w3_hi = smem_u32(smem + st * BF_STAGE_BYTES);
          const uint32_t w3_lo = w3_hi + BF_W1_BYTES;
          const uint32_t t_a = tmem_base + BF_COL_A + slot * 64;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t dbh = make_kmajor_sw128_desc(w3_hi + k * 32);
            const uint64_t dbl = make_kmajor_sw128_desc(w3_lo + k * 32);
            umma_tf32_ts(t_acc3, t_a + 32 + k * 8, dbh, idesc32, (c2 != 0 || k != 0) ? 1u : 0u);
            umma_tf32_ts(t_acc3, t_a + k * 8, dbl, idesc32, 1u);
            umma_tf32_ts(t_acc3, t_a + k * 8, dbh, idesc32, 1u);
          }
          umma_commit(&b_empty[st]);
          umma_commit(&a_empty[slot]);
          if (c2 == NC3 - 1) umma_commit(acc3_full);
        }
        __syncwarp();
      }
      gw += NC3, ga += NC3;
    }
  } else if (warp >= 4) {
    // ================= transform warps =================
    const int q = warp & 3;                    // TMEM lane quarter
    const int cw0 = ((warp - 4) >> 2) * BF_CW; // first of this warp's columns inside a 32-column chunk
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float beta1 = args.beta1 != nullptr ? __ldg(args.beta1) : 0.f;
    const float beta2 = args.beta2 != nullptr ? __ldg(args.beta2) : 0.f;
    const float nbl1 = -1.4426950408889634f * beta1, nbl2 = -1.4426950408889634f * beta2;
    const bool has_mul1 = args.mul1 != nullptr, has_mul2 = args.mul2 != nullptr;
    const int C = args.C;
    uint32_t ga = 0, gs = 0, it = 0;
    // the im2col row segment of the NEXT item is fetched while the current item's tail is processed
    float x0_next[BF_CW];
    auto fetch_x0 = [&](long long item) {
      const long long mm = (item / n_halves) * TC_BM + q * 32 + lane;
      if (item < num_items && mm < args.M) {
        ld8(args.x0 + mm * args.ldx + cw0, x0_next);
      } else {
#pragma unroll
        for (int j = 0; j < BF_CW; ++j) x0_next[j] = 0.f;
      }
    };
    fetch_x0(blockIdx.x);
    for (long long item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int nh = (int)(item % n_halves);
      const long long m = (item / n_halves) * TC_BM + q * 32 + lane;
      const bool valid = m < args.M;
      uint32_t hi[BF_CW], lo[BF_CW];
      float a[BF_CW];
      float mul[BF_CW], mul_next[BF_CW];      // multiplier rows are fetched one chunk ahead of their use
      if (has_mul1 && valid) ld8(args.mul1 + m * C + cw0, mul_next);
      // ---- X0 row -> TMEM (hi | lo)
      split8(x0_next, hi, lo);
      mbar_wait(x0_empty, (it & 1) ^ 1);
      tc_fence_after();
      tmem_st8(tmem_base + lane_base + BF_COL_X0 + cw0, hi);
      tmem_st8(tmem_base + lane_base + BF_COL_X0 + 32 + cw0, lo);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(x0_full);
      // ---- layer 1 chunks: S -> psi1 -> A ring
      for (int kc = 0; kc < NC; ++kc) {
        const int n0 = kc * TC_BK + cw0;
        if (has_mul1) {
#pragma unroll
          for (int j = 0; j < BF_CW; ++j) mul[j] = valid ? mul_next[j] : 0.f;
          if (valid && kc + 1 < NC) ld8(args.mul1 + m * C + n0 + TC_BK, mul_next);
        }
        if (kc == NC - 1 && has_mul2 && valid) ld8(args.mul2 + m * C + nh * 256 + cw0, mul_next);
        const uint32_t sc = gs + kc, buf = sc & 1;
        mbar_wait(&s_full[buf], (sc >> 1) & 1);
        tc_fence_after();
        uint32_t r[BF_CW];
        tmem_ld8(tmem_base + lane_base + BF_COL_S + buf * 32 + cw0, r);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[buf]);
        psi8<ACT>(r, a, args.bias1 != nullptr ? args.bias1 + n0 : nullptr, mul, has_mul1,
                  args.pre1_out != nullptr ? args.pre1_out + m * C + n0 : nullptr,
                  args.pre1_out != nullptr && valid && nh == 0, beta1, nbl1);
        split8(a, hi, lo);
        const uint32_t ac = ga + kc, slot = ac & 1;
        mbar_wait(&a_empty[slot], ((ac >> 1) & 1) ^ 1);
        tc_fence_after();
        tmem_st8(tmem_base + lane_base + BF_COL_A + slot * 64 + cw0, hi);
        tmem_st8(tmem_base + lane_base + BF_COL_A + slot * 64 + 32 + cw0, lo);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[slot]);
      }
      ga += NC, gs += NC;
      fetch_x0(item + gridDim.x);
      // ---- layer 2 accumulator: ACC -> psi2 -> A ring (layer-3 operand)
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      for (int c2 = 0; c2 < NC3; ++c2) {
        const int n0 = nh * 256 + c2 * TC_BK + cw0;
        if (has_mul2) {
#pragma unroll
          for (int j = 0; j < BF_CW; ++j) mul[j] = valid ? mul_next[j] : 0.f;
          if (valid && c2 + 1 < NC3) ld8(args.mul2 + m * C + n0 + TC_BK, mul_next);
        }
        uint32_t r[BF_CW];
        tmem_ld8(tmem_base + lane_base + BF_COL_ACC + c2 * TC_BK + cw0, r);
        psi8<ACT>(r, a, args.bias2 != nullptr ? args.bias2 + n0 : nullptr, mul, has_mul2,
                  args.pre2_out != nullptr ? args.pre2_out + m * C + n0 : nullptr,
                  args.pre2_out != nullptr && valid, beta2, nbl2);
        split8(a, hi, lo);
        const uint32_t ac = ga + c2, slot = ac & 1;
        mbar_wait(&a_empty[slot], ((ac >> 1) & 1) ^ 1);
        tc_fence_after();
        tmem_st8(tmem_base + lane_base + BF_COL_A + slot * 64 + cw0, hi);
        tmem_st8(tmem_base + lane_base + BF_COL_A + slot * 64 + 32 + cw0, lo);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[slot]);
      }
      ga += NC3;
      // ---- layer 3 accumulator -> out
      mbar_wait(acc3_full, it & 1);
      tc_fence_after();
      {
        uint32_t r[BF_CW];
        tmem_ld8(tmem_base + lane_base + BF_COL_S + 32 + cw0, r);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc3_empty);
        if (valid) {
          float* dst = args.out + m * args.ldo + cw0;
#pragma unroll
          for (int j = 0; j < BF_CW; ++j) {
            if (cw0 + j < args.N3) {
              if (n_halves > 1) {
                atomicAdd(dst + j, __uint_as_float(r[j]));    // exactly two addends on zeros: order-independent
              } else {
                dst[j] = __uint_as_float(r[j]);
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int ACT>
static int launch_branch3(const CUtensorMap* maps, const BranchArgs& a, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_branch3<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_BYTES) !=
        cudaSuccess) {
      set_error("branch3_tc: cannot set %d bytes of dynamic shared memory", BF_SMEM_BYTES);
      return -1;
    }
    attr_set = true;
  }
  const long long items = ((a.M + TC_BM - 1) / TC_BM) * (a.C / 256);
  const int grid = (int)(items < 148 ? items : 148);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(BF_THREADS);
  cfg.dynamicSmemBytes = BF_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = pdl_attr(&attr[0]);
  const cudaError_t err = cudaLaunchKernelEx(&cfg, k_branch3<ACT>, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], a);
  if (err != cudaSuccess) {
    set_error("k_branch3: launch failed: %s", cudaGetErrorString(err));
    return -1;
  }
  return check_launch("k_branch3");
}

}  // namespace impflow

using namespace impflow;

extern "C" int impflow_branch3_tc(const float* x0, long long ldx, const float* W1_hi, const float* W1_lo,
                                  const float* W2_hi, const float* W2_lo, const float* W3_hi, const float* W3_lo,
                                  const float* bias1, const float* bias2, const float* mul1, const float* mul2,
                                  float* pre1_out, float* pre2_out, float* out, long long ldo, long long M, int C,
                                  int N3, int act_kind, const float* beta1, const float* beta2, void* stream) {
  IMPFLOW_REQUIRE(M >= 1 && C >= 256 && N3 >= 1, "branch3_tc: empty problem M=%lld C=%d N3=%d", M, C, N3);
  if (C % 256 != 0 || N3 > 32 || ldx < 32 || (ldx % 4) != 0 || ldo < N3) {
    set_error("branch3_tc: needs C %% 256 == 0, N3 <= 32, ldx >= 32 and 16-byte aligned rows (C=%d N3=%d ldx=%lld)",
              C, N3, ldx);
    return -2;
  }
  const uintptr_t al = reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(W1_hi) |
                       reinterpret_cast<uintptr_t>(W1_lo) | reinterpret_cast<uintptr_t>(W2_hi) |
                       reinterpret_cast<uintptr_t>(W2_lo) | reinterpret_cast<uintptr_t>(W3_hi) |
                       reinterpret_cast<uintptr_t>(W3_lo) | reinterpret_cast<uintptr_t>(mul1) |
                       reinterpret_cast<uintptr_t>(mul2) | reinterpret_cast<uintptr_t>(pre1_out) |
                       reinterpret_cast<uintptr_t>(pre2_out) | reinterpret_cast<uintptr_t>(bias1) |
                       reinterpret_cast<uintptr_t>(bias2);
  if (al & 15) {
    set_error("branch3_tc: operand base pointers must be 16-byte aligned");
    return -2;
  }
  CUtensorMap maps[6];
  if (make_map(&maps[0], W1_hi, C, TC_BK, TC_BK, 32) || make_map(&maps[1], W1_lo, C, TC_BK, TC_BK, 32) ||
      make_map(&maps[2], W2_hi, C, C, C, 256) || make_map(&maps[3], W2_lo, C, C, C, 256) ||
      make_map(&maps[4], W3_hi, N3, C, C, 32) || make_map(&maps[5], W3_lo, N3, C, C, 32))
    return -1;
  BranchArgs a{x0, ldx, bias1, bias2, mul1, mul2, pre1_out, pre2_out, out, ldo, M, C, N3, beta1, beta2, g_gate};
  cudaStream_t s = (cudaStream_t)stream;
  switch (act_kind) {
    case IMPFLOW_ACT_LIPSWISH: return launch_branch3<IMPFLOW_ACT_LIPSWISH>(maps, a, s);
    case IMPFLOW_ACT_RELU: return launch_branch3<IMPFLOW_ACT_RELU>(maps, a, s);
    case IMPFLOW_ACT_SIN: return launch_branch3<IMPFLOW_ACT_SIN>(maps, a, s);
    default: return launch_branch3<IMPFLOW_ACT_NONE>(maps, a, s);
  }
}
