// Reverse-diffusion solver steps around the score networks: predictor-corrector sampler
// (solver.py:153-196) and S4 (solver.py:200-280).  One step = joint score evaluation(s) +
// per-graph norms -> batch means (-> optional 4-float all-reduce) -> fused state update.
// All per-step scalars come from the host-computed coefficient table (gdss_step_coef); the
// current row index lives in device memory so that one captured CUDA graph replays every step.
#include "gdss_common.cuh"
#include "gdss_internal.h"

namespace gdss {

#define LAUNCH_CHECK()                          \
  do {                                          \
    cudaError_t e__ = cudaGetLastError();       \
    if (e__ != cudaSuccess) return (int)e__;    \
  } while (0)

struct NoiseSrc {
  const float* inj_x;     // [iters][dx][B][N][F] or null
  const float* inj_adj;   // [iters][da][B][N][N] or null
  int dx, da;
  int dstride;            // Philox draw ids per step (4 unless a step has more draws: n_steps > 3 Langevin iterations)
  int first_step;
  uint64_t seed;
  int64_t graph_offset;
  int B, N, F;
};

// Philox draw id: the prior uses 0/1, step s draw d uses dstride*(s+1)+d  (dstride = 4 for every shipped configuration)
DEVINL uint32_t draw_id(const NoiseSrc& ns, int step, int d) { return (uint32_t)(ns.dstride * (step + 1) + d); }

// four raw N(0,1) for node-feature elements (i, 4*g .. 4*g+3) of graph b
DEVINL float4 noise_x4(const NoiseSrc& ns, int b, int step, int d, int i, int g) {
  if (ns.inj_x) {
    const float* r = ns.inj_x + (((int64_t)(step - ns.first_step) * ns.dx + d) * ns.B + b) * ns.N * ns.F + (int64_t)i * ns.F;
    float4 z;
    int f = 4 * g;
    z.x = f < ns.F ? r[f] : 0.f;
    z.y = f + 1 < ns.F ? r[f + 1] : 0.f;
    z.z = f + 2 < ns.F ? r[f + 2] : 0.f;
    z.w = f + 3 < ns.F ? r[f + 3] : 0.f;
    return z;
  }
  const int gw = (ns.F + 3) >> 2;
  return philox_normal4(ns.seed, (uint64_t)(ns.graph_offset + b), draw_id(ns, step, d), 0u, (uint32_t)(i * gw + g));
}

// four raw N(0,1) for adjacency elements (i, 4*g .. 4*g+3); only j > i is ever used
// (utils/graph_utils.py:80-83: triu(1) of a full N x N draw)
DEVINL float4 noise_a4(const NoiseSrc& ns, int b, int step, int d, int i, int g) {
  if (ns.inj_adj) {
    const float* r = ns.inj_adj + (((int64_t)(step - ns.first_step) * ns.da + d) * ns.B + b) * ns.N * ns.N + (int64_t)i * ns.N;
    float4 z;
    int j = 4 * g;
    z.x = j < ns.N ? r[j] : 0.f;
    z.y = j + 1 < ns.N ? r[j + 1] : 0.f;
    z.z = j + 2 < ns.N ? r[j + 2] : 0.f;
    z.w = j + 3 < ns.N ? r[j + 3] : 0.f;
    return z;
  }
  const int gw = (ns.N + 3) >> 2;
  return philox_normal4(ns.seed, (uint64_t)(ns.graph_offset + b), draw_id(ns, step, d), 1u, (uint32_t)(i * gw + g));
}

// ---------------------------------------------------------------------------------------
// prior (solver.py:174-178): x = mask_x(randn), adj = mask_adjs(triu(randn,1) + T)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prior_kernel(NoiseSrc ns, float* __restrict__ x, float* __restrict__ adj,
                                                    const float* __restrict__ flags) {
  const int b = blockIdx.x, N = ns.N, F = ns.F;
  const float* fl = flags + (int64_t)b * N;
  float* X = x + (int64_t)b * N * F;
  float* A = adj + (int64_t)b * N * N;
  const int gwx = (F + 3) >> 2, gwa = (N + 3) >> 2;
  for (int idx = threadIdx.x; idx < N * gwx; idx += 256) {
    int i = idx / gwx, g = idx - i * gwx;
    float4 z = noise_x4(ns, b, -1, 0, i, g);
    float f = fl[i];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (4 * g + q < F) X[i * F + 4 * g + q] = f4_get(z, q) * f;
  }
  for (int i = threadIdx.x; i < N; i += 256) A[i * N + i] = 0.f;
  for (int idx = threadIdx.x; idx < N * gwa; idx += 256) {
    int i = idx / gwa, g = idx - i * gwa;
    if (4 * g + 3 <= i) continue;
    float4 z = noise_a4(ns, b, -1, 1, i, g);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int j = 4 * g + q;
      if (j > i && j < N) {
        float v = f4_get(z, q) * fl[i] * fl[j];
        A[i * N + j] = v;
        A[j * N + i] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// norms_kernel: per-graph Frobenius norms of the scaled scores and of the masked corrector
// noise (solver.py:130-131, :141-142, :238-239, :250-251).  norms[b] = {|s_x|, |z_x|, |s_A|, |z_A|}
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) norms_kernel(NoiseSrc ns, const float* __restrict__ raw_x,
                                                    const float* __restrict__ raw_adj, const float* __restrict__ flags,
                                                    const int* __restrict__ neff, const gdss_step_coef* __restrict__ table,
                                                    const int* __restrict__ step_ptr, float* __restrict__ norms, int cd) {
  // cd: draw index of this corrector iteration (0 unless n_steps > 1)
  __shared__ float red[33];
  const int b = blockIdx.x, N = ns.N, F = ns.F, n = neff[b];
  const int step = *step_ptr;
  const gdss_step_coef c = table[step];
  const float* fl = flags + (int64_t)b * N;
  const float* RX = raw_x + (int64_t)b * N * F;
  const float* RA = raw_adj + (int64_t)b * N * N;
  const int gwx = (F + 3) >> 2, gwa = (N + 3) >> 2;
  float sx = 0.f, zx = 0.f, sa = 0.f, za = 0.f;
  for (int idx = threadIdx.x; idx < n * gwx; idx += 256) {
    int i = idx / gwx, g = idx - i * gwx;
    float4 z = noise_x4(ns, b, step, cd, i, g);
    float f = fl[i];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (4 * g + q < F) {
        float r = RX[i * F + 4 * g + q], zz = f4_get(z, q) * f;
        sx = fmaf(r, r, sx);
        zx = fmaf(zz, zz, zx);
      }
  }
  for (int idx = threadIdx.x; idx < n * gwa; idx += 256) {
    int i = idx / gwa, g = idx - i * gwa;
    if (4 * g + 3 <= i || 4 * g >= n) continue;
    float4 z = noise_a4(ns, b, step, cd, i, g);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int j = 4 * g + q;
      if (j > i && j < n) {
        float r = RA[i * N + j], zz = f4_get(z, q) * fl[i] * fl[j];
        sa = fmaf(r, r, sa);
        za = fmaf(zz, zz, za);
      }
    }
  }
  sx = block_sum(sx, red);
  zx = block_sum(zx, red);
  sa = block_sum(sa, red);
  za = block_sum(za, red);
  if (threadIdx.x == 0) {
    norms[4 * b + 0] = fabsf(c.cs[0]) * sqrtf(sx);
    norms[4 * b + 1] = sqrtf(zx);
    norms[4 * b + 2] = fabsf(c.cs[1]) * sqrtf(2.f * sa);   // both triangles
    norms[4 * b + 3] = sqrtf(2.f * za);
  }
}

// batch sums of the per-graph norms in a fixed order (deterministic): sums[0..3]
__global__ void __launch_bounds__(1024) sums_kernel(const float* __restrict__ norms, int B, float* __restrict__ sums) {
  __shared__ float red[33];
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  for (int b = threadIdx.x; b < B; b += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(norms + 4 * b);
    a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float s = block_sum(a[q], red);
    if (threadIdx.x == 0) sums[q] = s;
  }
}

// Langevin step size (solver.py:132): (snr * noise_norm / grad_norm)^2 * 2 * alpha
DEVINL float lang_eps(float snr, float nn, float gn, float alpha) {
  float t = snr * nn / gn;
  return t * t * 2.f * alpha;
}

enum { MODE_CORR = 0, MODE_PRED = 1, MODE_S4 = 2 };

struct UpdateArgs {
  NoiseSrc ns;
  float* x; float* adj; float* x_mean; float* adj_mean;
  const float* raw_x; const float* raw_adj;
  const float* flags; const int* neff;
  const gdss_step_coef* table; const int* step_ptr;
  const float* sums; float inv_gb;
  float* trace;          // [iters][8] or null
  int pred_draw;         // draw index of the predictor noise (n_steps after a corrector, else 0)
  int corr_draw;         // draw index of this Langevin iteration (0 .. n_steps-1)
};

// ---------------------------------------------------------------------------------------
// update_kernel<MODE>: fused elementwise state update of one solver phase.
//   CORR  solver.py:126-146   y += eps*s + sqrt(2 eps)*seps*z
//   PRED  solver.py:45-61 / :72-86 with the SDE algebra folded into (pa, pb, pc)
//   S4    solver.py:235-277   correction, half transition, drift, half transition
// grid = (chunks, B); the adjacency is updated on the strict upper triangle and mirrored.
// ---------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) update_kernel(UpdateArgs a) {
  const NoiseSrc& ns = a.ns;
  const int b = blockIdx.y, N = ns.N, F = ns.F;
  const int step = *a.step_ptr;
  const gdss_step_coef c = a.table[step];
  // The reference updates every element, padded nodes and the diagonal included (0 * coef + 0 * z): they
  // stay exactly 0 unless a coefficient is non-finite (e.g. the last S4 half-transition of a VP SDE with
  // few scales takes sqrt of a negative variance, sde.py:163-168) and then turn NaN everywhere.  Only in
  // that case is the full N x N range (diagonal included) walked; otherwise only the live n x n block.
  bool finite = true;
  {
    const float* cf = reinterpret_cast<const float*>(&c);
#pragma unroll
    for (int q = 0; q < 20; ++q) finite = finite && (fabsf(cf[q]) <= 3.0e38f);
  }
  float eps_x = 0.f, eps_a = 0.f;
  if (MODE != MODE_PRED) {
    const float gx = a.sums[0] * a.inv_gb, nx = a.sums[1] * a.inv_gb, ga = a.sums[2] * a.inv_gb, na = a.sums[3] * a.inv_gb;
    eps_x = lang_eps(c.snr, nx, gx, c.alpha[0]);
    eps_a = lang_eps(c.snr, na, ga, c.alpha[1]);
    if (a.trace && b == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
      float* tr = a.trace + 8 * (int64_t)(step - ns.first_step);
      tr[0] = eps_x; tr[1] = eps_a; tr[2] = gx; tr[3] = nx; tr[4] = ga; tr[5] = na; tr[6] = c.t; tr[7] = (float)step;
    }
  }
  if (MODE != MODE_PRED) finite = finite && (fabsf(eps_x) <= 3.0e38f) && (fabsf(eps_a) <= 3.0e38f);
  const int n = finite ? a.neff[b] : N;
  const int dg = finite ? 1 : 0;           // first column offset: strict upper triangle, or diagonal included
  const float nzx = sqrtf(eps_x * 2.f) * c.seps, nza = sqrtf(eps_a * 2.f) * c.seps;
  const float* fl = a.flags + (int64_t)b * N;
  const int gwx = (F + 3) >> 2, gwa = (N + 3) >> 2;
  const int stride = gridDim.x * 256, t0 = blockIdx.x * 256 + threadIdx.x;
  {
    float* X = a.x + (int64_t)b * N * F;
    float* XM = a.x_mean + (int64_t)b * N * F;
    const float* RX = a.raw_x + (int64_t)b * N * F;
    for (int idx = t0; idx < n * gwx; idx += stride) {
      int i = idx / gwx, g = idx - i * gwx;
      const float f = fl[i];
      float4 z0, z1, z2;
      if (MODE == MODE_CORR) z0 = noise_x4(ns, b, step, a.corr_draw, i, g);
      if (MODE == MODE_PRED) z0 = noise_x4(ns, b, step, a.pred_draw, i, g);
      if (MODE == MODE_S4) {
        z0 = noise_x4(ns, b, step, 0, i, g);
        z1 = noise_x4(ns, b, step, 1, i, g);
        z2 = noise_x4(ns, b, step, 2, i, g);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (4 * g + q >= F) continue;
        const int e = i * F + 4 * g + q;
        float y = X[e];
        const float r = RX[e];
        if (MODE == MODE_CORR) {
          y = y + eps_x * (c.cs[0] * r) + nzx * (f4_get(z0, q) * f);
        } else if (MODE == MODE_PRED) {
          float m = c.pa[0] * y + c.pb[0] * r;
          XM[e] = m;
          y = m + c.pc[0] * (f4_get(z0, q) * f);
        } else {
          y = y + eps_x * (c.cs[0] * r) + nzx * (f4_get(z0, q) * f);
          y = c.mu1[0] * y + c.sig1[0] * (f4_get(z1, q) * f);
          y = y + c.drift[0] * r;
          float m = c.mu2[0] * y;
          XM[e] = m;
          y = m + c.sig2[0] * (f4_get(z2, q) * f);
        }
        X[e] = y;
      }
    }
  }
  {
    float* A = a.adj + (int64_t)b * N * N;
    float* AM = a.adj_mean + (int64_t)b * N * N;
    const float* RA = a.raw_adj + (int64_t)b * N * N;
    for (int idx = t0; idx < n * gwa; idx += stride) {
      int i = idx / gwa, g = idx - i * gwa;
      if (4 * g + 3 < i + dg || 4 * g >= n) continue;
      const float fi = fl[i];
      float4 z0, z1, z2;
      if (MODE == MODE_CORR) z0 = noise_a4(ns, b, step, a.corr_draw, i, g);
      if (MODE == MODE_PRED) z0 = noise_a4(ns, b, step, a.pred_draw, i, g);
      if (MODE == MODE_S4) {
        z0 = noise_a4(ns, b, step, 0, i, g);
        z1 = noise_a4(ns, b, step, 1, i, g);
        z2 = noise_a4(ns, b, step, 2, i, g);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = 4 * g + q;
        if (j < i + dg || j >= n) continue;
        const float f = j == i ? 0.f : fi * fl[j];      // the symmetric noise has a zero diagonal (triu(1))
        const int e = i * N + j, et = j * N + i;
        float y = A[e];
        const float r = RA[e];
        if (MODE == MODE_CORR) {
          y = y + eps_a * (c.cs[1] * r) + nza * (f4_get(z0, q) * f);
        } else if (MODE == MODE_PRED) {
          float m = c.pa[1] * y + c.pb[1] * r;
          AM[e] = m;
          AM[et] = m;
          y = m + c.pc[1] * (f4_get(z0, q) * f);
        } else {
          y = y + eps_a * (c.cs[1] * r) + nza * (f4_get(z0, q) * f);
          y = c.mu1[1] * y + c.sig1[1] * (f4_get(z1, q) * f);
          y = y + c.drift[1] * r;
          float m = c.mu2[1] * y;
          AM[e] = m;
          AM[et] = m;
          y = m + c.sig2[1] * (f4_get(z2, q) * f);
        }
        A[e] = y;
        A[et] = y;
      }
    }
  }
}

__global__ void advance_kernel(int* step_ptr) { *step_ptr += 1; }
__global__ void set_step_kernel(int* step_ptr, int v) { *step_ptr = v; }

struct StepPlan {
  const gdss_ctx* ctx;
  UpdateArgs ua;
  int B;
  bool s4, langevin;
  char* ws;
  Workspace w;
  float* norms;
  float* sums;
  int* step_ptr;
  void* comm;
  int n_steps;           // Langevin iterations per step (solver.py:126, :138)
  float* x0; float* a0;  // n_steps > 1: the state the corrector phase started from
  int phases;            // GDSS_PHASE_CORRECTOR | GDSS_PHASE_PREDICTOR (gdss_solver_phase; gdss_sample runs both)
};

static int enqueue_step(const StepPlan& p, cudaStream_t st) {
  const UpdateArgs& u = p.ua;
  const int B = p.B;
  float* raw_x = const_cast<float*>(u.raw_x);
  float* raw_a = const_cast<float*>(u.raw_adj);
  const int N = u.ns.N, F = u.ns.F;
  int work = N * ((N + 3) / 4);
  int wx = N * ((F + 3) / 4);
  if (wx > work) work = wx;
  int chunks = (work + 255) / 256;
  if (chunks > 64) chunks = 64;
  dim3 ugrid(chunks, B);
  int rc = 0;
  // per-graph norms -> batch sums (-> all-reduce over the ranks) of the scores in raw_x / raw_a and of corrector draw cd
  auto norm_means = [&](int cd) -> int {
    norms_kernel<<<B, 256, 0, st>>>(u.ns, raw_x, raw_a, u.flags, u.neff, u.table, u.step_ptr, p.norms, cd);
    launched(K_NORMS, st);
    LAUNCH_CHECK();
    sums_kernel<<<1, 1024, 0, st>>>(p.norms, B, p.sums);
    launched(K_SUMS, st);
    LAUNCH_CHECK();
    if (p.comm) return nccl_allreduce_sum(p.comm, p.sums, 4, st);
    return 0;
  };
  if (p.s4) {
    rc = launch_eval(p.ctx, u.x, u.adj, u.flags, u.neff, raw_x, raw_a, B, p.ws, p.w, st, !p.ctx->fused);
    if (rc) return rc;
    rc = norm_means(0);
    if (rc) return rc;
    update_kernel<MODE_S4><<<ugrid, 256, 0, st>>>(u);
    launched(K_UPDATE, st);
    LAUNCH_CHECK();
  } else {
    if (p.langevin && (p.phases & GDSS_PHASE_CORRECTOR)) {
      const int K = p.n_steps;
      if (K > 1) {
        // solver.py:187-189: the x corrector iterates on (x_k, adj_0), the adjacency corrector on (x_0, adj_k)
        cudaError_t e = cudaMemcpyAsync(p.x0, u.x, sizeof(float) * (size_t)B * N * F, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(p.a0, u.adj, sizeof(float) * (size_t)B * N * N, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return (int)e;
      }
      for (int k = 0; k < K; ++k) {
        if (k == 0) {
          rc = launch_eval(p.ctx, u.x, u.adj, u.flags, u.neff, raw_x, raw_a, B, p.ws, p.w, st, !p.ctx->fused);
        } else {
          rc = launch_eval(p.ctx, u.x, p.a0, u.flags, u.neff, raw_x, nullptr, B, p.ws, p.w, st, !p.ctx->fused);
          if (rc) return rc;
          rc = launch_eval(p.ctx, p.x0, u.adj, u.flags, u.neff, nullptr, raw_a, B, p.ws, p.w, st, !p.ctx->fused);
        }
        if (rc) return rc;
        rc = norm_means(k);
        if (rc) return rc;
        UpdateArgs uk = u;
        uk.corr_draw = k;
        update_kernel<MODE_CORR><<<ugrid, 256, 0, st>>>(uk);
        launched(K_UPDATE, st);
        LAUNCH_CHECK();
      }
    }
    if (p.phases & GDSS_PHASE_PREDICTOR) {
      rc = launch_eval(p.ctx, u.x, u.adj, u.flags, u.neff, raw_x, raw_a, B, p.ws, p.w, st, !p.ctx->fused);
      if (rc) return rc;
      update_kernel<MODE_PRED><<<ugrid, 256, 0, st>>>(u);
      launched(K_UPDATE, st);
      LAUNCH_CHECK();
    }
  }
  advance_kernel<<<1, 1, 0, st>>>(p.step_ptr);
  launched(K_MISC, st);
  LAUNCH_CHECK();
  return 0;
}

}  // namespace gdss

using namespace gdss;

// phases / single_row / sums_out: gdss_solver_phase (one phase of ONE step: `table` then points at that step's row alone)
static int sample_impl(gdss_ctx* ctx, const gdss_sampler_desc* s, const gdss_step_coef* table, int32_t first_step,
                       int32_t n_iters, float* x, float* adj, const float* flags, float* x_mean, float* adj_mean,
                       int32_t B, int32_t init_prior, uint64_t seed, int64_t graph_offset, int64_t global_B,
                       const float* inj_x, const float* inj_adj, float* trace, void* nccl_comm, int32_t use_graph,
                       void* workspace, int64_t workspace_bytes, void* stream, int phases, bool single_row, float* sums_out) {
  if (!ctx || !s || !table || !x || !adj || !flags || !x_mean || !adj_mean || !workspace) return GDSS_E_BADARG;
  if (B <= 0 || n_iters < 0 || first_step < 0 || first_step + n_iters > GDSS_MAX_STEPS) return GDSS_E_BADARG;
  if (!ctx->has_x || !ctx->has_a) return GDSS_E_BADARG;
  if (s->n_steps < 1 || s->n_steps > 64) return GDSS_E_BADARG;     // (probability_flow lives in the table: pb halved, pc = 0)
  if (!nccl_comm && global_B != B) return GDSS_E_BADARG;
  if (init_prior && (inj_x || inj_adj)) return GDSS_E_BADARG;   // an injected run supplies its own prior
  if (global_B < B) return GDSS_E_BADARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return GDSS_E_UNSUPPORTED;   // no CPU fallback
  Workspace w;
  layout_workspace(ctx, B, &w);
  if (workspace_bytes < w.off[R_COUNT]) return GDSS_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const int N = ctx->N, F = ctx->F;
  const bool s4 = s->predictor == GDSS_PRED_S4;
  const bool langevin = !s4 && s->corrector == GDSS_CORR_LANGEVIN;

  StepPlan p;
  p.ctx = ctx;
  p.B = B;
  p.s4 = s4;
  p.langevin = langevin;
  p.ws = ws;
  p.w = w;
  p.norms = reinterpret_cast<float*>(ws + w.off[R_NORMS]);
  p.sums = reinterpret_cast<float*>(ws + w.off[R_MEANS]);
  p.step_ptr = reinterpret_cast<int*>(ws + w.off[R_STEP]);
  p.comm = nccl_comm;
  p.n_steps = langevin ? s->n_steps : 1;
  p.x0 = reinterpret_cast<float*>(ws + w.off[R_X0]);
  p.a0 = reinterpret_cast<float*>(ws + w.off[R_A0]);
  p.phases = phases;
  if (s4 && phases != (GDSS_PHASE_CORRECTOR | GDSS_PHASE_PREDICTOR)) return GDSS_E_BADARG;    // an S4 step is one unit (solver.py:229-278)
  int* neff = reinterpret_cast<int*>(ws + w.off[R_NEFF]);
  gdss_step_coef* dtable = reinterpret_cast<gdss_step_coef*>(ws + w.off[R_TABLE]);
  UpdateArgs& u = p.ua;
  u.ns.inj_x = inj_x;
  u.ns.inj_adj = inj_adj;
  u.ns.dx = u.ns.da = s4 ? 3 : (langevin ? p.n_steps + 1 : 1);
  u.ns.dstride = u.ns.dx > 4 ? u.ns.dx : 4;
  u.ns.first_step = first_step;
  u.ns.seed = seed;
  u.ns.graph_offset = graph_offset;
  u.ns.B = B;
  u.ns.N = N;
  u.ns.F = F;
  u.x = x; u.adj = adj; u.x_mean = x_mean; u.adj_mean = adj_mean;
  u.raw_x = reinterpret_cast<float*>(ws + w.off[R_RAW_X]);
  u.raw_adj = reinterpret_cast<float*>(ws + w.off[R_RAW_ADJ]);
  u.flags = flags; u.neff = neff;
  u.table = dtable; u.step_ptr = p.step_ptr;
  u.sums = p.sums; u.inv_gb = 1.0f / (float)global_B;
  u.trace = trace;
  u.pred_draw = langevin ? p.n_steps : 0;
  u.corr_draw = 0;

  cudaError_t e;
  if (single_row) e = cudaMemcpyAsync(dtable + first_step, table, sizeof(gdss_step_coef), cudaMemcpyHostToDevice, st);
  else e = cudaMemcpyAsync(dtable, table, sizeof(gdss_step_coef) * (size_t)(first_step + n_iters), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return (int)e;
  set_step_kernel<<<1, 1, 0, st>>>(p.step_ptr, first_step);
  launched(K_MISC, st);
  LAUNCH_CHECK();
  if ((int64_t)B * (N * (N - 1) / 2) >= (int64_t)1 << 30 || B >= (1 << 19)) return GDSS_E_UNSUPPORTED;
  int rc = launch_setup(ctx, flags, B, 1, ws, w, st);
  if (rc) return rc;
  if (init_prior) {
    prior_kernel<<<B, 256, 0, st>>>(u.ns, x, adj, flags);
    launched(K_MISC, st);
    LAUNCH_CHECK();
  }
  {
    // the means start as the state (solver.py returns x_mean of the last step; defined even for 0 steps),
    // which also defines the rows / columns of padded nodes that the update kernels never touch
    e = cudaMemcpyAsync(x_mean, x, sizeof(float) * (size_t)B * N * F, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyAsync(adj_mean, adj, sizeof(float) * (size_t)B * N * N, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return (int)e;
  }
  if (n_iters == 0) return GDSS_OK;
  if (ctx->fused) {
    // the fused evaluation writes the valid entries only: the score buffers are cleared once per call, not per step
    e = cudaMemsetAsync(const_cast<float*>(u.raw_x), 0, sizeof(float) * (size_t)B * N * F, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(const_cast<float*>(u.raw_adj), 0, sizeof(float) * (size_t)B * N * N, st);
    if (e != cudaSuccess) return (int)e;
  }

  // step 0 runs eagerly (sets kernel attributes, surfaces launch errors); the remaining steps replay
  // one captured graph -- the step index is read from device memory, so every replay is identical.
  rc = enqueue_step(p, st);
  if (rc) return rc;
  if (sums_out) {                            // the four batch sums of the corrector's norms (solver.py:130-132), device to device
    e = cudaMemcpyAsync(sums_out, p.sums, 4 * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return (int)e;
  }
  if (n_iters == 1) return GDSS_OK;
  if (!use_graph) {
    for (int it = 1; it < n_iters; ++it) {
      rc = enqueue_step(p, st);
      if (rc) return rc;
    }
    return GDSS_OK;
  }
  // Stream capture is not allowed on the legacy default stream: replay on a private blocking stream
  // (blocking streams are implicitly ordered with the legacy stream, so step 0 above is waited for).
  cudaStream_t cs = st;
  bool own_stream = false;
  if (st == nullptr || st == cudaStreamLegacy) {
    e = cudaStreamCreate(&cs);
    if (e != cudaSuccess) return (int)e;
    own_stream = true;
  }
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) {
    if (own_stream) cudaStreamDestroy(cs);
    return (int)e;
  }
  const int64_t before = gdss_launch_count();
  rc = enqueue_step(p, cs);
  const int64_t per_step = gdss_launch_count() - before;
  e = cudaStreamEndCapture(cs, &graph);
  if (rc == 0 && e != cudaSuccess) rc = (int)e;
  if (rc == 0) {
    e = cudaGraphInstantiate(&exec, graph, 0);
    if (e != cudaSuccess) rc = (int)e;
  }
  if (rc == 0) {
    for (int it = 1; it < n_iters; ++it) {
      e = cudaGraphLaunch(exec, cs);
      if (e != cudaSuccess) {
        rc = (int)e;
        break;
      }
    }
    count_launch((int)(per_step * (n_iters - 2)));      // the capture pass itself counted one step
    e = cudaStreamSynchronize(cs);
    if (rc == 0 && e != cudaSuccess) rc = (int)e;
  }
  if (exec) cudaGraphExecDestroy(exec);
  if (graph) cudaGraphDestroy(graph);
  if (own_stream) cudaStreamDestroy(cs);
  return rc;
}

extern "C" int gdss_sample(gdss_ctx* ctx, const gdss_sampler_desc* s, const gdss_step_coef* table, int32_t first_step,
                           int32_t n_iters, float* x, float* adj, const float* flags, float* x_mean, float* adj_mean,
                           int32_t B, int32_t init_prior, uint64_t seed, int64_t graph_offset, int64_t global_B,
                           const float* inj_x, const float* inj_adj, float* trace, void* nccl_comm, int32_t use_graph,
                           void* workspace, int64_t workspace_bytes, void* stream) {
  return sample_impl(ctx, s, table, first_step, n_iters, x, adj, flags, x_mean, adj_mean, B, init_prior, seed, graph_offset, global_B,
                     inj_x, inj_adj, trace, nccl_comm, use_graph, workspace, workspace_bytes, stream,
                     GDSS_PHASE_CORRECTOR | GDSS_PHASE_PREDICTOR, false, nullptr);
}

extern "C" int gdss_solver_phase(gdss_ctx* ctx, const gdss_sampler_desc* s, const gdss_step_coef* row, int32_t step, int32_t phases,
                                 float* x, float* adj, const float* flags, float* x_mean, float* adj_mean, int32_t B,
                                 uint64_t seed, int64_t graph_offset, int64_t global_B, const float* inj_x, const float* inj_adj,
                                 float* sums4_out, void* nccl_comm, void* workspace, int64_t workspace_bytes, void* stream) {
  if (phases < 1 || phases > (GDSS_PHASE_CORRECTOR | GDSS_PHASE_PREDICTOR) || step < 0 || step >= GDSS_MAX_STEPS) return GDSS_E_BADARG;
  return sample_impl(ctx, s, row, step, 1, x, adj, flags, x_mean, adj_mean, B, 0, seed, graph_offset, global_B, inj_x, inj_adj, nullptr,
                     nccl_comm, 0, workspace, workspace_bytes, stream, phases, true, sums4_out);
}
