// Micro-benchmark (not part of the library): issue cost of warp-level mma.sync m16n8k8 TF32 (the legacy tensor-core
// path) against packed FP32 FFMA2 on one SM sub-partition, and the accuracy of the 3xTF32 split through mma.sync.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void mma_rate(float* out, int iters, long long* cyc) {
  float c[NACC][4];
  uint32_t a[4], b[2];
  for (int q = 0; q < 4; ++q) a[q] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + q);
  b[0] = __float_as_uint(0.5f); b[1] = __float_as_uint(0.25f);
#pragma unroll
  for (int k = 0; k < NACC; ++k) for (int q = 0; q < 4; ++q) c[k][q] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < NACC; ++k) mma_tf32(c[k], a, b);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NACC; ++k) for (int q = 0; q < 4; ++q) s += c[k][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NACC>
__global__ void ffma2_rate(float* out, int iters, long long* cyc) {
  float2 c[NACC];
  float2 w = make_float2(1.0001f, 0.9999f);
  float a = 1.0f + threadIdx.x * 1e-6f;
#pragma unroll
  for (int k = 0; k < NACC; ++k) c[k] = make_float2(0.f, (float)k);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < NACC; ++k) c[k] = __ffma2_rn(make_float2(a, a), w, c[k]);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NACC; ++k) s += c[k].x + c[k].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__device__ __forceinline__ uint32_t tf32_hi(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}

// one warp: D[16x8] = A[16x8] B[8x8] by 3xTF32, compared on the host with fp64
__global__ void acc_probe(const float* A, const float* B, float* D, float* D1) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  float av[4] = {A[g * 8 + t], A[(g + 8) * 8 + t], A[g * 8 + t + 4], A[(g + 8) * 8 + t + 4]};
  float bv[2] = {B[t * 8 + g], B[(t + 4) * 8 + g]};
  uint32_t ah[4], al[4], bh[2], bl[2];
  for (int q = 0; q < 4; ++q) { ah[q] = tf32_hi(av[q]); al[q] = tf32_hi(av[q] - __uint_as_float(ah[q])); }
  for (int q = 0; q < 2; ++q) { bh[q] = tf32_hi(bv[q]); bl[q] = tf32_hi(bv[q] - __uint_as_float(bh[q])); }
  float c[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
  mma_tf32(c, al, bh);
  mma_tf32(c, ah, bl);
  mma_tf32(c, ah, bh);
  mma_tf32(c1, ah, bh);
  D[g * 8 + 2 * t] = c[0]; D[g * 8 + 2 * t + 1] = c[1]; D[(g + 8) * 8 + 2 * t] = c[2]; D[(g + 8) * 8 + 2 * t + 1] = c[3];
  D1[g * 8 + 2 * t] = c1[0]; D1[g * 8 + 2 * t + 1] = c1[1]; D1[(g + 8) * 8 + 2 * t] = c1[2]; D1[(g + 8) * 8 + 2 * t + 1] = c1[3];
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 8);
  const int iters = 4096;
  for (int warps = 1; warps <= 16; warps *= 2) {
    mma_rate<8><<<1, 32 * warps>>>(out, iters, cyc); cudaDeviceSynchronize();
    double m = (double)*cyc / (iters * 8.0);
    ffma2_rate<16><<<1, 32 * warps>>>(out, iters, cyc); cudaDeviceSynchronize();
    double f = (double)*cyc / (iters * 16.0);
    printf("warps/SM %2d (per sub-partition %.2f): cycles per mma.m16n8k8.tf32 per warp %.2f | per FFMA2 per warp %.2f\n", warps,
           warps / 4.0, m, f);
  }
  mma_rate<1><<<1, 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  printf("dependent mma latency %.1f cycles\n", (double)*cyc / iters);
  ffma2_rate<1><<<1, 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  printf("dependent FFMA2 latency %.1f cycles\n", (double)*cyc / iters);
  float hA[128], hB[64], *dA, *dB, *dD, *dD1, hD[128], hD1[128];
  srand(1);
  for (int i = 0; i < 128; ++i) hA[i] = (float)rand() / RAND_MAX * 2 - 1;
  for (int i = 0; i < 64; ++i) hB[i] = (float)rand() / RAND_MAX * 2 - 1;
  cudaMalloc(&dA, 512); cudaMalloc(&dB, 256); cudaMalloc(&dD, 512); cudaMalloc(&dD1, 512);
  cudaMemcpy(dA, hA, 512, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 256, cudaMemcpyHostToDevice);
  acc_probe<<<1, 32>>>(dA, dB, dD, dD1); cudaDeviceSynchronize();
  cudaMemcpy(hD, dD, 512, cudaMemcpyDeviceToHost); cudaMemcpy(hD1, dD1, 512, cudaMemcpyDeviceToHost);
  double e3 = 0, e1 = 0, ef = 0;
  for (int i = 0; i < 16; ++i) for (int j = 0; j < 8; ++j) {
    double r = 0; float rf = 0;
    for (int k = 0; k < 8; ++k) { r += (double)hA[i * 8 + k] * hB[k * 8 + j]; rf = fmaf(hA[i * 8 + k], hB[k * 8 + j], rf); }
    e3 = fmax(e3, fabs(hD[i * 8 + j] - r)); e1 = fmax(e1, fabs(hD1[i * 8 + j] - r)); ef = fmax(ef, fabs(rf - r));
  }
  printf("max abs err vs fp64: 3xTF32 %.3e, 1xTF32 %.3e, fp32 fma chain %.3e  (%s)\n", e3, e1, ef, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
