/* A C host of the plan-level ABI (include/se_b200.h): decodes a batch of waveforms with the CRN path using nothing but
 * libse_b200.so and the CUDA runtime -- no Python, no model code.  Built and driven by tests/test_gpu_models.py.
 *
 *   crn_plan_host <weights.bin> <wav_in.bin> <wav_out.bin> <B> <N> <use_graph>
 *
 * weights.bin: the fp32 tensors of the reference state-dict, concatenated in the order of se_crn_weights
 * (en_w[0..4], en_b[0..4], en_bn[0..4][0..3], lstm w_ih[0..1], w_hh, b_ih, b_hh, de_w[0..4], de_b[0..4], de_bn[0..4][0..3]).
 * wav_in.bin / wav_out.bin: [B][N] fp32.  This is what a maintainer of the reference would write around
 * CRN/crn_decode.py:17-68 if the decode loop were to be called from C. */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "se_b200.h"

static const int ENC_CH[6] = {1, 16, 32, 64, 128, 256};
static const int DEC_CI[5] = {512, 256, 128, 64, 32};
static const int DEC_CO[5] = {128, 64, 32, 16, 1};

static float* slurp(const char* path, size_t* n) {
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  float* p = (float*)malloc((size_t)sz);
  if (fread(p, 1, (size_t)sz, f) != (size_t)sz) return NULL;
  fclose(f);
  *n = (size_t)sz / sizeof(float);
  return p;
}

int main(int argc, char** argv) {
  if (argc != 7) return 2;
  const int B = atoi(argv[4]), N = atoi(argv[5]), use_graph = atoi(argv[6]);
  size_t nw = 0, nx = 0;
  float* w = slurp(argv[1], &nw);
  float* x = slurp(argv[2], &nx);
  if (!w || !x || nx != (size_t)B * N) return 3;
  se_crn_weights cw;
  const float* p = w;
#define TAKE(dst, count) do { (dst) = p; p += (count); } while (0)
  for (int i = 0; i < 5; ++i) TAKE(cw.en_w[i], (size_t)ENC_CH[i + 1] * ENC_CH[i] * 6);
  for (int i = 0; i < 5; ++i) TAKE(cw.en_b[i], ENC_CH[i + 1]);
  for (int i = 0; i < 5; ++i) for (int j = 0; j < 4; ++j) TAKE(cw.en_bn[i][j], ENC_CH[i + 1]);
  for (int l = 0; l < 2; ++l) TAKE(cw.lstm_w_ih[l], (size_t)4096 * 1024);
  for (int l = 0; l < 2; ++l) TAKE(cw.lstm_w_hh[l], (size_t)4096 * 1024);
  for (int l = 0; l < 2; ++l) TAKE(cw.lstm_b_ih[l], 4096);
  for (int l = 0; l < 2; ++l) TAKE(cw.lstm_b_hh[l], 4096);
  for (int i = 0; i < 5; ++i) TAKE(cw.de_w[i], (size_t)DEC_CI[i] * DEC_CO[i] * 6);
  for (int i = 0; i < 5; ++i) TAKE(cw.de_b[i], DEC_CO[i]);
  for (int i = 0; i < 5; ++i) for (int j = 0; j < 4; ++j) TAKE(cw.de_bn[i][j], DEC_CO[i]);
  if ((size_t)(p - w) != nw) { fprintf(stderr, "weights.bin has %zu floats, expected %zu\n", nw, (size_t)(p - w)); return 4; }

  se_plan_t* plan = NULL;
  if (se_plan_create_crn(&cw, B, N, &plan) != SE_OK) { fprintf(stderr, "%s\n", se_last_error()); return 5; }
  se_plan_set_graph(plan, use_graph);
  float *dx = NULL, *dy = NULL;
  cudaMalloc((void**)&dx, (size_t)B * N * 4);
  cudaMalloc((void**)&dy, (size_t)B * N * 4);
  cudaMemcpy(dx, x, (size_t)B * N * 4, cudaMemcpyHostToDevice);
  cudaStream_t s;
  cudaStreamCreate(&s);
  for (int rep = 0; rep < 3; ++rep)       /* with use_graph: capture on the first call, replay afterwards */
    if (se_enhance_crn(plan, dx, N, dy, N, B, N, NULL, 1.0f, s) != SE_OK) { fprintf(stderr, "%s\n", se_last_error()); return 6; }
  if (cudaStreamSynchronize(s) != cudaSuccess) { fprintf(stderr, "%s\n", cudaGetErrorString(cudaGetLastError())); return 7; }
  cudaMemcpy(x, dy, (size_t)B * N * 4, cudaMemcpyDeviceToHost);
  FILE* f = fopen(argv[3], "wb");
  fwrite(x, 4, (size_t)B * N, f);
  fclose(f);
  printf("workspace_bytes %lld launches %llu\n", se_query_workspace(plan), se_launch_count());
  se_plan_destroy(plan);
  return 0;
}
