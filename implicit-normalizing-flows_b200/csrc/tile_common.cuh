// Device helpers shared by the fused tile kernels (branch_fused.cu: narrow -> C -> C -> narrow with 9c <= 32;
// chain23_fused.cu: C -> C -> narrow for the wider scales): TS-form tcgen05.mma (A operand in tensor memory),
// 8-column tensor-memory loads / stores of the transform warps, the tf32 hi/lo split and the psi epilogue.
#pragma once
#include "tc_common.cuh"

namespace impflow {

constexpr int BF_XF_WARPS = 16;                // transform warps: 4 per TMEM lane quarter
constexpr int BF_CW = 32 / (BF_XF_WARPS / 4);  // columns of a 32-column chunk per transform warp

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_c),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tf32 hi/lo split: round-to-nearest, ties away from zero (what cvt.rna.tf32.f32 does for finite values),
// written as two integer ops so that the transform loop stays short
__device__ __forceinline__ void split8(const float* v, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int j = 0; j < BF_CW; ++j) {
    const uint32_t hb = (__float_as_uint(v[j]) + 0x1000u) & 0xffffe000u;
    hi[j] = hb;
    lo[j] = __float_as_uint(v[j] - __uint_as_float(hb));
  }
}
__device__ __forceinline__ void ld8(const float* p, float* v) {
  const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int j = 0; j < BF_CW / 4; ++j) {
    const float4 t = __ldg(p4 + j);
    v[4 * j] = t.x, v[4 * j + 1] = t.y, v[4 * j + 2] = t.z, v[4 * j + 3] = t.w;
  }
}
__device__ __forceinline__ void st8(float* p, const float* v) {
  float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int j = 0; j < BF_CW / 4; ++j) p4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

// psi: the accumulator values of one row segment -> the next layer's operand values.
//   multiplier mode: a = acc * mul;   activation mode: v = acc + bias (optionally stored), a = act(v).
// LipSwish x*sigmoid(beta*x)/1.1 runs on the SFU fast paths with the constants folded:
//   e = 2^(x * (-beta*log2 e)),  a = x * rcp(1.1 + 1.1 e)       (5 instructions per element)
template <int ACT>
__device__ __forceinline__ void psi8(const uint32_t* r, float* a, const float* bias, const float* mul, bool has_mul,
                                     float* pre_out, bool store_pre, float beta, float neg_beta_log2e) {
  if (has_mul) {
#pragma unroll
    for (int j = 0; j < BF_CW; ++j) a[j] = __uint_as_float(r[j]) * mul[j];
    return;
  }
  float v[BF_CW];
  if (bias != nullptr) {
    float b[BF_CW];
    ld8(bias, b);
#pragma unroll
    for (int j = 0; j < BF_CW; ++j) v[j] = __uint_as_float(r[j]) + b[j];
  } else {
#pragma unroll
    for (int j = 0; j < BF_CW; ++j) v[j] = __uint_as_float(r[j]);
  }
  if (store_pre) st8(pre_out, v);
#pragma unroll
  for (int j = 0; j < BF_CW; ++j) {
    if (ACT == IMPFLOW_ACT_LIPSWISH) {
      float e, s;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v[j] * neg_beta_log2e));
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(fmaf(e, 1.1f, 1.1f)));
      a[j] = v[j] * s;
    } else {
      a[j] = act_eval<ACT>(v[j], 0, beta);
    }
  }
}

}  // namespace impflow
