// Stand-alone bring-up probe for the tcgen05 building blocks the final-edge-MLP kernel uses
// (not part of the library): TMEM alloc, tcgen05.st of an A operand, kind::tf32 MMA with A in TMEM
// and B in shared memory (K-major, no swizzle), tcgen05.commit -> mbarrier, tcgen05.ld of D, and the
// 3xTF32 split (hi*hi + lo*hi + hi*lo) against an fp64 host product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && ./umma_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define M 128
#define NOUT 96
#define K1 48

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_bdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
  return d;                 // layout_type 0 = no swizzle, base_offset 0
}

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ bool mbar_wait(uint32_t mbar, uint32_t parity) {
  for (int it = 0; it < 4000000; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(mbar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}

// mode 0: single tf32 MMA chain (operands already exactly representable) ; mode 1: 3xTF32 split
__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D,
                                             int mode, int* status) {
  extern __shared__ __align__(128) unsigned char smraw[];
  float* Bhi = reinterpret_cast<float*>(smraw);            // [K1/4][NOUT][4]
  float* Blo = Bhi + K1 * NOUT;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  // B operand: W[n][k] (nn.Linear layout, K contiguous) -> [k/4][n][k%4]
  for (int idx = tid; idx < NOUT * K1; idx += 128) {
    int n = idx / K1, k = idx % K1;
    float w = W[idx];
    float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    float lo = w - hi;
    int o = (k >> 2) * (NOUT * 4) + n * 4 + (k & 3);
    Bhi[o] = mode ? hi : w;
    Blo[o] = lo;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // make the generic-proxy smem writes of B visible to the async proxy (tensor core reads)
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tb = tmem_base_s;
  const uint32_t lane_base = tb + ((uint32_t)(warp * 32) << 16);
  const uint32_t colAhi = 0, colAlo = K1, colD = 2 * K1;      // columns: Ahi [0,48) Alo [48,96) D [96,192)
  // A operand: thread = row
  {
    const float* a = A + (size_t)tid * K1;
    for (int c0 = 0; c0 < K1; c0 += 8) {
      float hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float v = a[c0 + q];
        float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        hi[q] = mode ? h : v;
        lo[q] = v - h;
      }
      tmem_st8(lane_base + colAhi + c0, hi);
      tmem_st8(lane_base + colAlo + c0, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NOUT >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (tid == 0) {
    const uint32_t lbo = NOUT * 16, sbo = 128;
    uint32_t acc = 0;
    const int passes = mode ? 3 : 1;
    for (int ps = 0; ps < passes; ++ps) {
      const uint32_t acol = (ps == 1) ? colAlo : colAhi;
      const float* bsrc = (ps == 2) ? Blo : Bhi;
      for (int s = 0; s < K1 / 8; ++s) {
        uint64_t bd = make_bdesc(smem_u32(bsrc) + s * 2 * lbo, lbo, sbo);
        mma_tf32_ts(tb + colD, tb + acol + s * 8, bd, idesc, acc);
        acc = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar)) : "memory");
  }
  bool ok = mbar_wait(smem_u32(&mbar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (!ok) {
    if (tid == 0) *status = 1;
  } else {
    for (int c0 = 0; c0 < NOUT; c0 += 8) {
      float v[8];
      tmem_ld8(lane_base + colD + c0, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) D[(size_t)tid * NOUT + c0 + q] = v[q];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tb) : "memory");
}

int main() {
  float *hA = (float*)malloc(M * K1 * 4), *hW = (float*)malloc(NOUT * K1 * 4), *hD = (float*)malloc(M * NOUT * 4);
  float *dA, *dW, *dD;
  int* dS;
  cudaMalloc(&dA, M * K1 * 4);
  cudaMalloc(&dW, NOUT * K1 * 4);
  cudaMalloc(&dD, M * NOUT * 4);
  cudaMalloc(&dS, 4);
  size_t smem = 2 * K1 * NOUT * 4;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int fail = 0;
  for (int mode = 0; mode < 2; ++mode) {
    srand(7 + mode);
    for (int i = 0; i < M * K1; ++i) hA[i] = mode ? (float)rand() / RAND_MAX * 4.f - 2.f : (float)(rand() % 9 - 4);
    for (int i = 0; i < NOUT * K1; ++i) hW[i] = mode ? (float)rand() / RAND_MAX * 2.f - 1.f : (float)(rand() % 5 - 2);
    cudaMemcpy(dA, hA, M * K1 * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, hW, NOUT * K1 * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, M * NOUT * 4);
    cudaMemset(dS, 0, 4);
    probe<<<1, 128, smem>>>(dA, dW, dD, mode, dS);
    cudaError_t e = cudaDeviceSynchronize();
    int st = -1;
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hD, dD, M * NOUT * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0, sq = 0, sqr = 0;
    int bad_r = -1, bad_c = -1;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < NOUT; ++n) {
        double r = 0;
        for (int k = 0; k < K1; ++k) r += (double)hA[m * K1 + k] * (double)hW[n * K1 + k];
        double d = fabs(r - (double)hD[m * NOUT + n]);
        if (!(d <= maxerr)) { maxerr = d; bad_r = m; bad_c = n; }
        if (fabs(r) > maxref) maxref = fabs(r);
        sq += d * d;
        sqr += r * r;
      }
    printf("mode %d: cuda=%s status=%d max_abs_err=%.3e (at row %d col %d) max_ref=%.3e rel_l2=%.3e  D[0][0..3]=%g %g %g %g\n", mode,
           cudaGetErrorString(e), st, maxerr, bad_r, bad_c, maxref, sqrt(sq / sqr), hD[0], hD[1], hD[2], hD[3]);
    if (e != cudaSuccess || st != 0) fail = 1;
    if (mode == 0 && maxerr != 0.0) fail = 1;
    if (mode == 1 && !(sqrt(sq / sqr) < 2e-6)) fail = 1;
  }
  printf(fail ? "PROBE FAIL\n" : "PROBE OK\n");
  return fail;
}
