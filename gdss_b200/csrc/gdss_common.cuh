// Common device helpers for the GDSS sampling kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DEVINL __device__ __forceinline__

namespace gdss {

// ---- activations ---------------------------------------------------------------------
// ELU (alpha = 1) as torch.nn.functional.elu (models/layers.py:150).  ex2.approx based:
// absolute error ~1e-7 on the negative branch, exact on the positive one.
// Direct MUFU forms (no denormal fix-up code around them: a flushed result only meets exp(-88) ~ 0):
DEVINL float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
DEVINL float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
DEVINL float elu_f(float v) { return v > 0.f ? v : (ex2_ftz(v * 1.4426950408889634f) - 1.f); }

// tanh via one ex2.approx and one rcp.approx: |err| <= ~3e-7 absolute (tanh.approx.f32 is
// 2^-11 and breaks the 1e-4 per-step parity bar).
DEVINL float tanh_f(float v) {
  v = fminf(fmaxf(v, -15.f), 15.f);
  float e = __expf(2.f * v);
  return 1.f - __fdividef(2.f, e + 1.f);
}

// ---- reductions ----------------------------------------------------------------------
DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum, result valid in every thread.  `red` = >= 33 floats of shared memory.
// Deterministic: fixed shuffle tree, fixed warp order.
DEVINL float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// ---- Philox4x32-10 counter-based RNG -> N(0,1) ---------------------------------------
// Counter layout (documented in DESIGN.md): c = (group, graph_lo, graph_hi | domain<<31, draw)
// key = (seed_lo, seed_hi).  One call yields four normals for elements 4*group .. 4*group+3
// of one (graph, draw).  Results are independent of the launch geometry and of how the batch
// is sharded over GPUs.
DEVINL uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

DEVINL void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;   // (0,1]
  const float u2 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.f * __logf(fminf(u1, 0.99999994f)));
  float s, c;
  sincospif(2.f * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// four N(0,1) for elements [4*group, 4*group+4) of (graph, draw) in domain 0 (x) / 1 (adj)
DEVINL float4 philox_normal4(uint64_t seed, uint64_t graph, uint32_t draw, uint32_t domain, uint32_t group) {
  uint4 c = make_uint4(group, (uint32_t)graph, ((uint32_t)(graph >> 32) & 0x7fffffffu) | (domain << 31), draw);
  uint2 k = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  uint4 r = philox4x32_10(c, k);
  float4 z;
  box_muller(r.x, r.y, z.x, z.y);
  box_muller(r.z, r.w, z.z, z.w);
  return z;
}

DEVINL float f4_get(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

}  // namespace gdss
