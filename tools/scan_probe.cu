// Stand-alone timing of candidate inner loops of the window search (sm_100a): the same tables in
// shared memory (TL[k][j], VD[k][8], Df[k]), the same per-thread work (one steering rate j, eight
// accelerations), different ways of getting sin/cos of the eight headings and of packing the
// position / cost arithmetic.  Prints clocks per loop iteration (= 8 hypothesis-steps per thread,
// 256 per warp) per SM sub-partition, for a given number of CTAs per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/_build/scan_probe tools/scan_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int kN = 60;     // steps per window
constexpr int kGS = 256;   // steering rates (threads per CTA)

__device__ __forceinline__ float2 pk(float a, float b) { return make_float2(a, b); }

struct Out { float J[8]; float stat; };

// V0: generic scan, 16 MUFU
__device__ __forceinline__ void loop_generic(const float* tl, const float* vd, const float2* Df, Out& o) {
  float th[8], ex[8], ey[8], J[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) th[c] = ex[c] = ey[c] = J[c] = 0.f;
  float tv = 0.f, tlmax = 0.f, vmax = 0.f;
#pragma unroll 1
  for (int k = 1; k <= kN; ++k) {
    const float tlk = tl[(k - 1) * kGS];
    const float4 va = *reinterpret_cast<const float4*>(vd + (k - 1) * 8);
    const float4 vb = *reinterpret_cast<const float4*>(vd + (k - 1) * 8 + 4);
    const float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
    const float2 d = Df[k];
    tv = fmaf(v[7], fabsf(tlk), tv);
    tlmax = fmaxf(tlmax, fabsf(tlk));
    vmax = fmaxf(vmax, v[7]);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      th[c] = fmaf(v[c], tlk, th[c]);
      float sn, cs;
      __sincosf(th[c], &sn, &cs);
      ex[c] = fmaf(v[c], cs, ex[c] - d.x);
      ey[c] = fmaf(v[c], sn, ey[c] - d.y);
      J[c] = fmaf(ex[c], ex[c], J[c]);
      J[c] = fmaf(ey[c], ey[c], J[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) o.J[c] = J[c];
  o.stat = tv + tlmax + vmax;
}

struct Aff { float vwdt, dt2, a0, da, amax; };

// how the sin/cos of the eight headings A + a_c B are produced
//   TRIG 0: 8 direct MUFU pairs (16 MUFU)
//   TRIG 1: current scheme, MUFU at {0,1,4,5} + one rotation by 2 da B (10 MUFU, 4 rotations)
//   TRIG 2: MUFU at 3, chain of rotations by +-da B (4 MUFU, 7 rotations, depth 4)
//   TRIG 3: MUFU at 3 and for da B, 2 da B (6 MUFU, 7 rotations, depth 2)
//   TRIG 4: MUFU at {1, 5} and for da B, 2 da B: 0=1-d 2=1+d 3=1+2d | 4=5-d 6=5+d 7=5+2d  (8 MUFU, 6 rotations, depth 1)
__device__ __forceinline__ void rot(float c, float s, float cr, float sr, float& co, float& so) {
  co = fmaf(c, cr, -(s * sr));
  so = fmaf(s, cr, c * sr);
}

template <int TRIG>
__device__ __forceinline__ void trig8(float A, float B, const Aff& f, float* cs, float* sn) {
  if (TRIG == 0) {
#pragma unroll
    for (int c = 0; c < 8; ++c) __sincosf(fmaf(fmaf((float)c, f.da, f.a0), B, A), &sn[c], &cs[c]);
  } else if (TRIG == 1) {
    const float a1 = f.a0 + f.da, a4 = fmaf(4.f, f.da, f.a0), a5 = fmaf(5.f, f.da, f.a0);
    float sr, cr;
    __sincosf(fmaf(f.a0, B, A), &sn[0], &cs[0]);
    __sincosf(fmaf(a1, B, A), &sn[1], &cs[1]);
    __sincosf(fmaf(a4, B, A), &sn[4], &cs[4]);
    __sincosf(fmaf(a5, B, A), &sn[5], &cs[5]);
    __sincosf((f.da + f.da) * B, &sr, &cr);
    rot(cs[0], sn[0], cr, sr, cs[2], sn[2]);
    rot(cs[1], sn[1], cr, sr, cs[3], sn[3]);
    rot(cs[4], sn[4], cr, sr, cs[6], sn[6]);
    rot(cs[5], sn[5], cr, sr, cs[7], sn[7]);
  } else if (TRIG == 2) {
    const float a3 = fmaf(3.f, f.da, f.a0);
    float sr, cr;
    __sincosf(fmaf(a3, B, A), &sn[3], &cs[3]);
    __sincosf(f.da * B, &sr, &cr);
    rot(cs[3], sn[3], cr, sr, cs[4], sn[4]);
    rot(cs[3], sn[3], cr, -sr, cs[2], sn[2]);
    rot(cs[4], sn[4], cr, sr, cs[5], sn[5]);
    rot(cs[2], sn[2], cr, -sr, cs[1], sn[1]);
    rot(cs[5], sn[5], cr, sr, cs[6], sn[6]);
    rot(cs[1], sn[1], cr, -sr, cs[0], sn[0]);
    rot(cs[6], sn[6], cr, sr, cs[7], sn[7]);
  } else if (TRIG == 3) {
    const float a3 = fmaf(3.f, f.da, f.a0);
    float sr, cr, sr2, cr2;
    __sincosf(fmaf(a3, B, A), &sn[3], &cs[3]);
    const float x = f.da * B;
    __sincosf(x, &sr, &cr);
    __sincosf(x + x, &sr2, &cr2);
    rot(cs[3], sn[3], cr, sr, cs[4], sn[4]);
    rot(cs[3], sn[3], cr, -sr, cs[2], sn[2]);
    rot(cs[3], sn[3], cr2, sr2, cs[5], sn[5]);
    rot(cs[3], sn[3], cr2, -sr2, cs[1], sn[1]);
    rot(cs[4], sn[4], cr2, sr2, cs[6], sn[6]);
    rot(cs[2], sn[2], cr2, -sr2, cs[0], sn[0]);
    rot(cs[5], sn[5], cr2, sr2, cs[7], sn[7]);
  } else if (TRIG == 5) {
    // MUFU at {1, 2, 5, 6}; the outer four by the three-term recurrence c(i-1) = 2 cos(da B) c(i) - c(i+1)
    // (one FFMA per value instead of a rotation's two; error 5 eps_MUFU instead of 2): 9 MUFU
    const float a1 = f.a0 + f.da, a2 = fmaf(2.f, f.da, f.a0), a5 = fmaf(5.f, f.da, f.a0), a6 = fmaf(6.f, f.da, f.a0);
    __sincosf(fmaf(a1, B, A), &sn[1], &cs[1]);
    __sincosf(fmaf(a2, B, A), &sn[2], &cs[2]);
    __sincosf(fmaf(a5, B, A), &sn[5], &cs[5]);
    __sincosf(fmaf(a6, B, A), &sn[6], &cs[6]);
    float K = __cosf(f.da * B);
    K += K;
    cs[0] = fmaf(K, cs[1], -cs[2]); sn[0] = fmaf(K, sn[1], -sn[2]);
    cs[3] = fmaf(K, cs[2], -cs[1]); sn[3] = fmaf(K, sn[2], -sn[1]);
    cs[4] = fmaf(K, cs[5], -cs[6]); sn[4] = fmaf(K, sn[5], -sn[6]);
    cs[7] = fmaf(K, cs[6], -cs[5]); sn[7] = fmaf(K, sn[6], -sn[5]);
  } else {
    const float a1 = f.a0 + f.da, a5 = fmaf(5.f, f.da, f.a0);
    float sr, cr, sr2, cr2;
    __sincosf(fmaf(a1, B, A), &sn[1], &cs[1]);
    __sincosf(fmaf(a5, B, A), &sn[5], &cs[5]);
    const float x = f.da * B;
    __sincosf(x, &sr, &cr);
    __sincosf(x + x, &sr2, &cr2);
    rot(cs[1], sn[1], cr, -sr, cs[0], sn[0]);
    rot(cs[1], sn[1], cr, sr, cs[2], sn[2]);
    rot(cs[1], sn[1], cr2, sr2, cs[3], sn[3]);
    rot(cs[5], sn[5], cr, -sr, cs[4], sn[4]);
    rot(cs[5], sn[5], cr, sr, cs[6], sn[6]);
    rot(cs[5], sn[5], cr2, sr2, cs[7], sn[7]);
  }
}

// PACK 0: scalar position / cost; 1: packed FFMA2 / FADD2 (pairs (0,1) (2,3) ...); 2: pairs 0,1 packed, 2,3 scalar
// VDREG 1: the step lengths max(0, vwdt + a_c k dt^2) worked out in registers instead of read from VD[k][8]
template <int TRIG, int PACK, int JSPLIT, int VDREG = 0>
__device__ __forceinline__ void loop_affine(const float* tl, const float* vd, const float2* Df, const Aff& f,
                                            Out& o) {
  float A = 0.f, B = 0.f, tv = 0.f, tlmax = 0.f, vmax = 0.f, kdt2 = 0.f;
  float ex[8], ey[8], J[8], J2[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) ex[c] = ey[c] = J[c] = J2[c] = 0.f;
#pragma unroll 1
  for (int k = 1; k <= kN; ++k) {
    const float tlk = tl[(k - 1) * kGS];
    float v[8];
    kdt2 += f.dt2;
    if (VDREG) {
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = fmaxf(fmaf(fmaf((float)c, f.da, f.a0), kdt2, f.vwdt), 0.f);
    } else {
      const float4 va = *reinterpret_cast<const float4*>(vd + (k - 1) * 8);
      const float4 vb = *reinterpret_cast<const float4*>(vd + (k - 1) * 8 + 4);
      v[0] = va.x; v[1] = va.y; v[2] = va.z; v[3] = va.w; v[4] = vb.x; v[5] = vb.y; v[6] = vb.z; v[7] = vb.w;
    }
    const float2 d = Df[k];
    A = fmaf(f.vwdt, tlk, A);
    B = fmaf(kdt2, tlk, B);
    tv = fmaf(fmaf(f.amax, kdt2, f.vwdt), fabsf(tlk), tv);
    tlmax = fmaxf(tlmax, fabsf(tlk));
    vmax = fmaxf(vmax, v[7]);
    float cs[8], sn[8];
    trig8<TRIG>(A, B, f, cs, sn);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const bool packed = PACK == 1 || (PACK == 2 && q < 2);
      if (packed) {
        float2 e2x = pk(ex[2 * q], ex[2 * q + 1]), e2y = pk(ey[2 * q], ey[2 * q + 1]);
        float2 j2 = pk(J[2 * q], J[2 * q + 1]), j2b = pk(J2[2 * q], J2[2 * q + 1]);
        const float2 v2 = pk(v[2 * q], v[2 * q + 1]);
        e2x = __ffma2_rn(v2, pk(cs[2 * q], cs[2 * q + 1]), __fadd2_rn(e2x, pk(-d.x, -d.x)));
        e2y = __ffma2_rn(v2, pk(sn[2 * q], sn[2 * q + 1]), __fadd2_rn(e2y, pk(-d.y, -d.y)));
        if (JSPLIT) {
          j2 = __ffma2_rn(e2x, e2x, j2);
          j2b = __ffma2_rn(e2y, e2y, j2b);
        } else {
          j2 = __ffma2_rn(e2x, e2x, j2);
          j2 = __ffma2_rn(e2y, e2y, j2);
        }
        ex[2 * q] = e2x.x; ex[2 * q + 1] = e2x.y; ey[2 * q] = e2y.x; ey[2 * q + 1] = e2y.y;
        J[2 * q] = j2.x; J[2 * q + 1] = j2.y; J2[2 * q] = j2b.x; J2[2 * q + 1] = j2b.y;
      } else {
#pragma unroll
        for (int c = 2 * q; c < 2 * q + 2; ++c) {
          ex[c] = fmaf(v[c], cs[c], ex[c] - d.x);
          ey[c] = fmaf(v[c], sn[c], ey[c] - d.y);
          J[c] = fmaf(ex[c], ex[c], J[c]);
          if (JSPLIT) J2[c] = fmaf(ey[c], ey[c], J2[c]);
          else J[c] = fmaf(ey[c], ey[c], J[c]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) o.J[c] = J[c] + J2[c];
  o.stat = tv + tlmax + vmax;
}

template <int VAR, int MINB>
__global__ void __launch_bounds__(kGS, MINB) probe(int windows, float* sink) {
  extern __shared__ __align__(16) float smem[];
  float* TL = smem;                       // [kN][kGS]
  float* VD = TL + kN * kGS;              // [kN][8]
  float2* Df = reinterpret_cast<float2*>(VD + kN * 8);   // [kN + 1]
  const int j = threadIdx.x;
  for (int i = threadIdx.x; i < kN * kGS; i += blockDim.x)
    TL[i] = 0.02f * sinf(0.01f * (i % kGS) + 0.001f * (i / kGS));
  for (int i = threadIdx.x; i < kN * 8; i += blockDim.x) VD[i] = 0.4f + 0.002f * (i % 8) + 0.0005f * (i / 8);
  for (int i = threadIdx.x; i <= kN; i += blockDim.x) Df[i] = make_float2(0.4f, 0.01f * i);
  __syncthreads();
  Aff f{0.4f, 0.0025f, -1.f + 0.01f * blockIdx.x, 0.08f, 1.f};
  float acc = 0.f;
  for (int w = 0; w < windows; ++w) {
    Out o;
    f.a0 += 1e-4f;
    if (VAR == 0) loop_generic(TL + j, VD, Df, o);
    if (VAR == 1) loop_affine<1, 1, 1>(TL + j, VD, Df, f, o);   // the shipped fast scan
    if (VAR == 2) loop_affine<1, 0, 0>(TL + j, VD, Df, f, o);   // same trig, scalar arithmetic
    if (VAR == 3) loop_affine<2, 0, 0>(TL + j, VD, Df, f, o);   // 4 MUFU chain, scalar
    if (VAR == 4) loop_affine<2, 1, 0>(TL + j, VD, Df, f, o);   // 4 MUFU chain, packed pos/cost
    if (VAR == 5) loop_affine<3, 0, 0>(TL + j, VD, Df, f, o);   // 6 MUFU depth 2, scalar
    if (VAR == 6) loop_affine<3, 1, 0>(TL + j, VD, Df, f, o);   // 6 MUFU depth 2, packed
    if (VAR == 7) loop_affine<4, 0, 0>(TL + j, VD, Df, f, o);   // 8 MUFU depth 1, scalar
    if (VAR == 8) loop_affine<4, 1, 0>(TL + j, VD, Df, f, o);   // 8 MUFU depth 1, packed
    if (VAR == 9) loop_affine<4, 2, 0>(TL + j, VD, Df, f, o);   // 8 MUFU depth 1, half packed
    if (VAR == 10) loop_affine<2, 2, 0>(TL + j, VD, Df, f, o);  // 4 MUFU chain, half packed
    if (VAR == 11) loop_affine<1, 2, 0>(TL + j, VD, Df, f, o);  // 10 MUFU, half packed
    if (VAR == 12) loop_affine<0, 0, 0>(TL + j, VD, Df, f, o);  // affine headings, 16 MUFU, scalar
    if (VAR == 13) loop_affine<0, 1, 0>(TL + j, VD, Df, f, o);  // affine headings, 16 MUFU, packed
    if (VAR == 14) loop_affine<1, 1, 1, 1>(TL + j, VD, Df, f, o);  // the shipped scan without the VD table
    if (VAR == 15) loop_affine<5, 1, 1>(TL + j, VD, Df, f, o);     // 9 MUFU, three-term recurrence, packed, split J
#pragma unroll
    for (int c = 0; c < 8; ++c) acc += o.J[c];
    acc += o.stat;
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int VAR, int MINB>
static void run(const char* name, int sms) {
  const int smem = (kN * kGS + kN * 8) * 4 + (kN + 1) * 8;
  auto kern = probe<VAR, MINB>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kGS, smem);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  if (per_sm > MINB) per_sm = MINB;
  const int blocks = sms * per_sm, windows = 64;
  float* sink;
  cudaMalloc(&sink, sizeof(float) * blocks * kGS);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(a);
    kern<<<blocks, kGS, smem>>>(windows, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (r) best = ms < best ? ms : best;
  }
  int khz;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double cycles = best * 1e-3 * khz * 1e3;
  const double warp_iters_per_smsp = (double)per_sm * (kGS / 32) * windows * kN / 4.0;
  const double hs = (double)blocks * kGS * 8.0 * windows * kN;
  printf("%-44s regs %3d spill %3zu  CTAs/SM %d  %7.3f ms  %6.1f clk/iter/SMSP  %6.3f T hyp-steps/s\n", name,
         fa.numRegs, (size_t)fa.localSizeBytes, per_sm, best, cycles / warp_iters_per_smsp,
         hs / (best * 1e-3) / 1e12);
  cudaFree(sink);
}

#define RUN3(V, NAME)            \
  run<V, 2>(NAME, sms);          \
  run<V, 3>(NAME, sms);          \
  run<V, 4>(NAME, sms);

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  RUN3(0, "generic 16 MUFU scalar");
  RUN3(12, "affine 16 MUFU scalar");
  RUN3(13, "affine 16 MUFU packed");
  RUN3(1, "shipped fast: 10 MUFU, packed, split J");
  RUN3(14, "shipped, step lengths in registers (no VD)");
  RUN3(15, "9 MUFU three-term recurrence, packed");
  RUN3(2, "10 MUFU scalar");
  RUN3(11, "10 MUFU half packed");
  RUN3(7, "8 MUFU depth1 scalar");
  RUN3(8, "8 MUFU depth1 packed");
  RUN3(9, "8 MUFU depth1 half packed");
  RUN3(5, "6 MUFU depth2 scalar");
  RUN3(6, "6 MUFU depth2 packed");
  RUN3(3, "4 MUFU chain scalar");
  RUN3(4, "4 MUFU chain packed");
  RUN3(10, "4 MUFU chain half packed");
  return 0;
}
