// Shared helpers for the se_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/se_b200.h"

namespace se {

// ---- error plumbing (C-ABI returns int status; message kept per thread) -------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError -> SE_ERR_CUDA
bool profiler_attached();             // Nsight Compute / compute-sanitizer injected into this process (api.cu)

#define SE_REQUIRE(cond, ...)                    \
  do {                                           \
    if (!(cond)) {                               \
      se::set_error(__VA_ARGS__);                \
      return SE_ERR_SHAPE;                       \
    }                                            \
  } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---- activations -------------------------------------------------------------------
// Accurate (not fast-math) forms: parity with the fp32 reference is the first gate.
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float elu_f(float x) { return x > 0.0f ? x : expm1f(x); }
// torch.nn.Softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

__device__ __forceinline__ float apply_act(float x, int act, float param = 0.0f) {
  switch (act) {
    case SE_ACT_PRELU: return x >= 0.0f ? x : param * x;
    case SE_ACT_ELU: return elu_f(x);
    case SE_ACT_SOFTPLUS: return softplus_f(x);
    case SE_ACT_RELU: return fmaxf(x, 0.0f);
    case SE_ACT_SIGMOID: return sigmoid_f(x);
    case SE_ACT_TANH: return tanhf(x);
    default: return x;
  }
}

// fast-math gates: ex2.approx + approximate divide, |error| ~ 2e-7 (fp32 rounding class); the accurate expf/tanhf/IEEE
// divide of common.cuh cost 2800 cycles of a 20000-cycle step here (profiles/lstm_tc_phases_v1_r01.json)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  const float t = __expf(-2.0f * fabsf(x));
  return copysignf(__fdividef(1.0f - t, 1.0f + t), x);
}

// ELU for the tensor-core epilogues: expm1f (about 30 instructions with two branches) was the largest single item of the
// conv epilogue, and the epilogue -- eight warps, two per scheduler, dependent ALU chains -- is what bounds the small-
// channel conv layers (profiles/ncu_conv_f16_r02p_*: 70 instructions per output element, issue slots 33 % busy, tensor
// pipe 17-49 %).  ex2.approx - 1 has an ABSOLUTE error of ~1e-7 (fp32 rounding of a value near 1); next to zero, where
// that would be a large relative error, the second-order series is exact to 1.6e-10.
__device__ __forceinline__ float fast_elu(float x) {
  const float e = __expf(x) - 1.0f;
  const float s = fmaf(0.5f * x, x, x);
  return x > 0.0f ? x : (x > -9.765625e-4f ? s : e);
}

// ---- packed fp32x2 FMA (Blackwell FFMA2) -----------------------------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(SE_NO_FFMA2)
  return __ffma2_rn(a, b, c);
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

// ---- async copy helpers --------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- mbarrier + 1-D bulk (TMA engine) copy -------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
// cp.async.bulk global -> shared (UBLKCP in SASS).  dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
               "l"(gmem_src), "r"(bytes), "r"(b)
               : "memory");
}

}  // namespace se
