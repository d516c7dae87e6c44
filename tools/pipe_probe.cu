// Issue-rate microbenchmarks for the pipes the scan loop uses (sm_100a): prints warp-instructions
// per clock per SM sub-partition for FFMA, FFMA2, FADD2, FMUL2, MUFU and two mixes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/pipe_probe tools/pipe_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int KIND>
__global__ void probe(int iters, float* sink) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (KIND == 0) {  // FFMA, 16 chains
    float a[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) a[q] = 1.f + 1e-3f * ((tid + q) & 255);
    float m = 0.9999f + 1e-9f * tid, b = 1e-4f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 16; ++q) a[q] = fmaf(a[q], m, b);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) acc += a[q];
  } else if (KIND == 1 || KIND == 2 || KIND == 3) {  // FFMA2 / FADD2 / FMUL2, 8 packed chains
    float2 a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = make_float2(1.f + 1e-3f * ((tid + q) & 255), 1.f + 2e-3f * (tid & 63));
    float2 m = make_float2(0.9999f + 1e-9f * tid, 0.99991f), b = make_float2(1e-4f, 2e-4f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (KIND == 1) a[q] = __ffma2_rn(a[q], m, b);
        if (KIND == 2) a[q] = __fadd2_rn(a[q], b);
        if (KIND == 3) a[q] = __fmul2_rn(a[q], m);
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += a[q].x + a[q].y;
  } else if (KIND == 4) {  // MUFU sin + cos, 8 chains (16 MUFU, 8 FMUL.RZ, 8 FFMA per iteration)
    float a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = 0.001f * (tid & 1023) + 0.37f * q;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = fmaf(__cosf(a[q]), 0.5f, __sinf(a[q]));
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += a[q];
  } else if (KIND == 5) {  // the scalar scan mix: per hypothesis 2 MUFU + 1 FMUL.RZ + 5 FFMA + 2 FADD
    float th[8], x[8], y[8], J[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { th[q] = 0.001f * (tid & 1023) + 0.1f * q; x[q] = y[q] = J[q] = 0.f; }
    float v = 0.5f + 1e-6f * tid, tl = 0.01f, dx = 0.4f, dy = 0.01f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        th[q] = fmaf(v, tl, th[q]);
        float s, c;
        __sincosf(th[q], &s, &c);
        x[q] = fmaf(v, c, x[q] - dx);
        y[q] = fmaf(v, s, y[q] - dy);
        J[q] = fmaf(x[q], x[q], J[q]);
        J[q] = fmaf(y[q], y[q], J[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += J[q];
  } else if (KIND == 6) {  // packed mix without MUFU: per pair FADD2 x2, FFMA2 x4
    float2 x[4], y[4], Jx[4], Jy[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) x[q] = y[q] = Jx[q] = Jy[q] = make_float2(0.f, 0.f);
    float2 v = make_float2(0.5f + 1e-6f * tid, 0.6f), c = make_float2(0.8f, 0.7f), s = make_float2(0.6f, 0.71f);
    float2 nd = make_float2(-0.4f, -0.4f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        x[q] = __ffma2_rn(v, c, __fadd2_rn(x[q], nd));
        y[q] = __ffma2_rn(v, s, __fadd2_rn(y[q], nd));
        Jx[q] = __ffma2_rn(x[q], x[q], Jx[q]);
        Jy[q] = __ffma2_rn(y[q], y[q], Jy[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) acc += Jx[q].x + Jx[q].y + Jy[q].x + Jy[q].y;
  } else if (KIND == 7) {  // IMAD, 16 chains
    int a[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) a[q] = tid + q;
    int m = 3 + (tid & 1), b = 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 16; ++q) a[q] = a[q] * m + b;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) acc += (float)a[q];
  }
  sink[tid] = acc;
}

template <int KIND>
static void run(const char* name, double inst_per_iter, int sms) {
  const int blocks = sms * 4, threads = 256, iters = 2048;
  float* sink;
  cudaMalloc(&sink, sizeof(float) * blocks * threads);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(a);
    probe<KIND><<<blocks, threads>>>(iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (r) best = ms < best ? ms : best;
  }
  int khz;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double warps = (double)blocks * threads / 32, cycles = best * 1e-3 * khz * 1e3;
  const double per_smsp = warps * iters * inst_per_iter / (cycles * sms * 4);
  printf("%-28s %8.3f ms  %6.3f warp-inst/clk/SMSP (at %d MHz nominal)  -> %5.2f clk per inst\n", name, best,
         per_smsp, khz / 1000, 1.0 / per_smsp);
  cudaFree(sink);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  run<0>("FFMA (3-reg)", 16, sms);
  run<1>("FFMA2", 8, sms);
  run<2>("FADD2", 8, sms);
  run<3>("FMUL2", 8, sms);
  run<4>("MUFU (counting MUFU only)", 16, sms);
  run<5>("scalar scan mix (80 inst)", 80, sms);
  run<6>("packed mix (24 inst)", 24, sms);
  run<7>("IMAD", 16, sms);
  return 0;
}
