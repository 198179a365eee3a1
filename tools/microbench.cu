// Pipe-rate microbenchmarks on sm_100a (development tool): FFMA, FFMA2, mma.sync tf32 / bf16.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench.cu -o gpurun_out/microbench
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__global__ void k_ffma(float* out, float a, float b) {
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// distinct register operands per FMA (GEMM-like: acc += a_i * b_j)
__global__ void k_ffma_outer(float* out, const float* in) {
  float a[4], b[4], acc[16];
  for (int i = 0; i < 4; ++i) { a[i] = in[i + threadIdx.x % 3]; b[i] = in[8 + i + threadIdx.x % 5]; }
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i * 4 + j] = fmaf(a[i], b[j], acc[i * 4 + j]);
    a[0] += 1e-9f;  // keep the loop from being hoisted
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2_outer(float* out, const float* in) {
  float2 a[4], b[2], acc[8];
  for (int i = 0; i < 4; ++i) { float v = in[i + threadIdx.x % 3]; a[i] = make_float2(v, v); }
  for (int i = 0; i < 2; ++i) b[i] = make_float2(in[8 + i + threadIdx.x % 5], in[12 + i]);
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(0.f, 0.f);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) acc[i * 2 + j] = __ffma2_rn(a[i], b[j], acc[i * 2 + j]);
    a[0].x += 1e-9f;
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mma_tf32(float* out, const float* in) {
  unsigned a[4], b[2];
  float d[8][4];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(in[i]);
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(in[4 + i]);
  for (int t = 0; t < 8; ++t) for (int i = 0; i < 4; ++i) d[t][i] = 0.f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int t = 0; t < 8; ++t)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(d[t][0]), "+f"(d[t][1]), "+f"(d[t][2]), "+f"(d[t][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
  for (int t = 0; t < 8; ++t) for (int i = 0; i < 4; ++i) s += d[t][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mma_bf16(float* out, const float* in) {
  unsigned a[4], b[2];
  float d[8][4];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(in[i]);
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(in[4 + i]);
  for (int t = 0; t < 8; ++t) for (int i = 0; i < 4; ++i) d[t][i] = 0.f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int t = 0; t < 8; ++t)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(d[t][0]), "+f"(d[t][1]), "+f"(d[t][2]), "+f"(d[t][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
  for (int t = 0; t < 8; ++t) for (int i = 0; i < 4; ++i) s += d[t][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float run(F f, const char* name, double ops_per_thread, int threads = 256, int blocks = 148 * 4) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  double total = ops_per_thread * threads * blocks;
  printf("%-28s %8.3f ms  %8.2f T(op)/s  (err=%s)\n", name, ms, total / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  return ms;
}

int main() {
  float *out, *in;
  cudaMalloc(&out, 148 * 4 * 256 * 4 * 4);
  cudaMalloc(&in, 64 * 4);
  float h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0f + i * 1e-3f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const int B = 148 * 4, T = 256;
  run([&] { k_ffma<<<B, T>>>(out, 1.0001f, 1e-7f); }, "FFMA (2-src same regs) FLOP", 2.0 * 16 * ITERS);
  run([&] { k_ffma_outer<<<B, T>>>(out, in); }, "FFMA outer-product FLOP", 2.0 * 16 * ITERS);
  run([&] { k_ffma2_outer<<<B, T>>>(out, in); }, "FFMA2 outer-product FLOP", 2.0 * 16 * ITERS);
  // per warp: 8 MMAs x (16*8*8) MAC; per thread share = /32
  run([&] { k_mma_tf32<<<B, T>>>(out, in); }, "mma.sync m16n8k8 tf32 FLOP", 2.0 * 8 * 16 * 8 * 8 / 32 * ITERS);
  run([&] { k_mma_bf16<<<B, T>>>(out, in); }, "mma.sync m16n8k16 bf16 FLOP", 2.0 * 8 * 16 * 8 * 16 / 32 * ITERS);
  return 0;
}
