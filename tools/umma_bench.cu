// tcgen05.mma issue-rate microbenchmark (sm_100a): cycles per MMA for the operand/shape mixes the LSTM
// recurrence (csrc/lstm_tc.cu) and the 3xTF32 GEMM (csrc/gemm_tc.cu) use.  One CTA per SM, operands are
// whatever is in shared / tensor memory (values do not matter for timing).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I<pkg>/csrc tools/umma_bench.cu -o tools/umma_bench.bin
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace se;

__device__ __forceinline__ void umma_ts(unsigned d, unsigned a, uint64_t b, unsigned idesc, unsigned acc, int kind) {
  if (kind == 0)
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_ss(unsigned d, uint64_t a, uint64_t b, unsigned idesc, unsigned acc, int kind) {
  if (kind == 0)
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
constexpr unsigned idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}

// mode: 0 SS same D | 1 TS same D | 2 TS alternating D0/D1 | 3 recurrence mix (TS D0, TS D1, SS D1)
//       4 GEMM mix (3 SS into one D) | 5 recurrence mix reordered (all D0 of a k-block, then all D1)
template <bool ELECT>
__global__ void __launch_bounds__(128, 1) bench(int mode, int N, int kind, int reps, long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ unsigned slot;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 0.001f * (i & 255);
  asm volatile("fence.proxy.async;\n" ::: "memory");
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tm = slot;
  if (ELECT ? (threadIdx.x < 32 && elect_one()) : (threadIdx.x == 0)) {
    const unsigned idesc = kind == 0 ? make_idesc_tf32(128, N) : idesc_bf16(128, N);
    const uint64_t da = make_smem_desc(base), db = make_smem_desc(base + 32768), db2 = make_smem_desc(base + 65536);
    const unsigned d0 = tm + 256, d1 = tm + 384;   // N <= 128 columns each (N = 256: d0 only)
    const long long c0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adv = (uint64_t)((k * 32) >> 4);
        const unsigned at = tm + (unsigned)(k * 8);
        if (mode == 0) {
          umma_ss(d0, da + adv, db + adv, idesc, 1, kind);
        } else if (mode == 1) {
          umma_ts(d0, at, db + adv, idesc, 1, kind);
        } else if (mode == 2) {
          umma_ts((k & 1) ? d1 : d0, at, db + adv, idesc, 1, kind);
        } else if (mode == 3) {
          umma_ts(d0, at, db + adv, idesc, 1, kind);
          umma_ts(d1, at, db2 + adv, idesc, 1, kind);
          umma_ss(d1, da + adv, db + adv, idesc, 1, kind);
        } else if (mode == 4) {
          umma_ss(d0, da + adv, db + adv, idesc, 1, kind);
          umma_ss(d0, da + adv, db2 + adv, idesc, 1, kind);
          umma_ss(d0, da + adv, db + adv, idesc, 1, kind);
        }
      }
      if (mode == 5) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts(d0, tm + (unsigned)(k * 8), db + (uint64_t)((k * 32) >> 4), idesc, 1, kind);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts(d1, tm + (unsigned)(k * 8), db2 + (uint64_t)((k * 32) >> 4), idesc, 1, kind);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(d1, da + (uint64_t)((k * 32) >> 4), db + (uint64_t)((k * 32) >> 4), idesc, 1, kind);
      }
    }
    umma_commit(&bar);
    unsigned it = 0;
    unsigned ok = 0;
    while (!ok) {
      asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}"
                   : "=r"(ok)
                   : "r"(smem_u32(&bar))
                   : "memory");
      if (++it > (1u << 24)) break;
    }
    const long long c1 = clock64();
    if (blockIdx.x == 0) {
      out[0] = c1 - c0;
      out[1] = ok;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(bench<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"SS same D", "TS same D", "TS alt D0/D1", "recurrence mix TS,TS,SS", "GEMM mix SS,SS,SS",
                         "recurrence mix grouped by D"};
  const int reps = 512;
  for (int el = 0; el < 2; ++el)
  for (int kind = 0; kind < 2; ++kind)
    for (int N : {64, 128, 256})
      for (int mode = 0; mode < 6; ++mode) {
        if (N == 256 && (mode == 2 || mode == 3 || mode == 5)) continue;
        const int per = (mode == 3 || mode == 4 || mode == 5) ? 12 : 4;
        if (el)
          bench<true><<<148, 128, smem>>>(mode, N, kind, reps, d);
        else
          bench<false><<<148, 128, smem>>>(mode, N, kind, reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
          printf("%s N=%d mode %d: %s\n", kind ? "bf16" : "tf32", N, mode, cudaGetErrorString(e));
          return 1;
        }
        const double cyc = (double)h[0] / (reps * per);
        const double k_per = kind ? 16 : 8;
        printf("%s %s M=128 N=%3d %-30s %7.1f cyc/MMA  (floor M*N/256 = %3d)  %6.1f TFLOP/s/chip@1.85GHz ok=%lld\n",
               el ? "elect " : "lane==0", kind ? "bf16" : "tf32", N, names[mode], cyc, 128 * N / 256, 2.0 * 128 * N * k_per / cyc * 1.85e9 * 148 / 1e12,
               h[1]);
      }
  return 0;
}
