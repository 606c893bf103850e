// Denoising-score-matching loss, forward only (losses.py:37-81 with likelihood_weighting = False): what the reference's
// trainer evaluates on the test loader every epoch under torch.no_grad() (trainer.py:84-99), and the forward half of a
// training step.  The per-graph SDE scalars (marginal mean coefficient, std; losses.py:51,56 via sde.marginal_prob) are
// computed by the host mirror with the reference's own torch expressions and passed as coef[B][6]:
//   mean_x, std_x, div_x, mean_adj, std_adj, div_adj      (div = std of that SDE for VP / subVP: score = -net / std,
//                                                          losses.py:19; 0 for VE: score = net, losses.py:26)
#include "gdss_common.cuh"
#include "gdss_internal.h"

namespace gdss {

#define LAUNCH_CHECK()                          \
  do {                                          \
    cudaError_t e__ = cudaGetLastError();       \
    if (e__ != cudaSuccess) return (int)e__;    \
  } while (0)

// flags = node_flags(adj) (graph_utils.py:33-39); z = gen_noise (graph_utils.py:78-86) from the supplied raw draws;
// perturbed = mask(mean * y + std * z) (losses.py:50-58)
__global__ void __launch_bounds__(256) dsm_perturb_kernel(const float* __restrict__ x, const float* __restrict__ adj,
                                                          const float* __restrict__ coef, const float* __restrict__ raw_x,
                                                          const float* __restrict__ raw_adj, float* __restrict__ flags,
                                                          float* __restrict__ px, float* __restrict__ padj, float* __restrict__ zx,
                                                          float* __restrict__ zadj, int N, int F) {
  extern __shared__ float fl[];
  const int b = blockIdx.x;
  const float* A = adj + (int64_t)b * N * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < N; ++j) s += fabsf(A[i * N + j]);
    const float f = s > 1e-5f ? 1.f : 0.f;
    fl[i] = f;
    flags[(int64_t)b * N + i] = f;
  }
  __syncthreads();
  const float mx = coef[b * 6 + 0], sx = coef[b * 6 + 1], ma = coef[b * 6 + 3], sa = coef[b * 6 + 4];
  for (int idx = threadIdx.x; idx < N * F; idx += blockDim.x) {
    const int i = idx / F;
    const int64_t e = (int64_t)b * N * F + idx;
    const float z = raw_x[e] * fl[i];
    zx[e] = z;
    px[e] = (mx * x[e] + sx * z) * fl[i];
  }
  const float* R = raw_adj + (int64_t)b * N * N;
  for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
    const int i = idx / N, j = idx - i * N;
    const int64_t e = (int64_t)b * N * N + idx;
    float z = i < j ? R[i * N + j] : (i > j ? R[j * N + i] : 0.f);       // z.triu(1) + transpose
    z = z * fl[i] * fl[j];
    zadj[e] = z;
    padj[e] = (ma * A[idx] + sa * z) * fl[i] * fl[j];
  }
}

// per graph: reduce_op(square(score * std + z)) for both networks (losses.py:63-67)
__global__ void __launch_bounds__(256) dsm_loss_kernel(const float* __restrict__ net_x, const float* __restrict__ net_adj,
                                                       const float* __restrict__ zx, const float* __restrict__ zadj,
                                                       const float* __restrict__ coef, float* __restrict__ lossb, int N, int F,
                                                       int reduce_mean) {
  __shared__ float red[40];
  const int b = blockIdx.x;
  const float sx = coef[b * 6 + 1], dx = coef[b * 6 + 2], sa = coef[b * 6 + 4], da = coef[b * 6 + 5];
  float ax = 0.f, aa = 0.f;
  for (int idx = threadIdx.x; idx < N * F; idx += blockDim.x) {
    const int64_t e = (int64_t)b * N * F + idx;
    const float s = dx != 0.f ? -net_x[e] / dx : net_x[e];
    const float v = s * sx + zx[e];
    ax = fmaf(v, v, ax);
  }
  for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
    const int64_t e = (int64_t)b * N * N + idx;
    const float s = da != 0.f ? -net_adj[e] / da : net_adj[e];
    const float v = s * sa + zadj[e];
    aa = fmaf(v, v, aa);
  }
  ax = block_sum(ax, red);
  aa = block_sum(aa, red);
  if (threadIdx.x == 0) {
    lossb[b * 2 + 0] = reduce_mean ? ax / (float)(N * F) : 0.5f * ax;
    lossb[b * 2 + 1] = reduce_mean ? aa / (float)(N * N) : 0.5f * aa;
  }
}

// torch.mean over the batch (losses.py:79), fixed summation order
__global__ void __launch_bounds__(1024) dsm_mean_kernel(const float* __restrict__ lossb, int B, float* __restrict__ out) {
  __shared__ float red[40];
  float a = 0.f, c = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    a += lossb[b * 2 + 0];
    c += lossb[b * 2 + 1];
  }
  a = block_sum(a, red);
  c = block_sum(c, red);
  if (threadIdx.x == 0) {
    out[0] = a / (float)B;
    out[1] = c / (float)B;
  }
}

int launch_dsm_perturb(const float* x, const float* adj, const float* coef, const float* raw_x, const float* raw_adj, float* flags,
                       float* px, float* padj, float* zx, float* zadj, int B, int N, int F, cudaStream_t st) {
  dsm_perturb_kernel<<<B, 256, sizeof(float) * N, st>>>(x, adj, coef, raw_x, raw_adj, flags, px, padj, zx, zadj, N, F);
  launched(K_MISC, st);
  LAUNCH_CHECK();
  return 0;
}

}  // namespace gdss

extern "C" int64_t gdss_dsm_scratch_floats(const gdss_ctx* ctx, int32_t B) {
  if (!ctx || B <= 0) return GDSS_E_BADARG;
  const int64_t N = ctx->N, F = ctx->F;
  return (int64_t)B * (N + 3 * N * F + 3 * N * N + 2);
}

extern "C" int gdss_dsm_loss(gdss_ctx* ctx, const float* x, const float* adj, const float* coef, const float* raw_x,
                             const float* raw_adj, int32_t B, int32_t reduce_mean, float* scratch, int64_t scratch_floats,
                             float* loss_out, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace gdss;
  if (!ctx || !x || !adj || !coef || !raw_x || !raw_adj || !scratch || !loss_out || !workspace || B <= 0) return GDSS_E_BADARG;
  if (!ctx->has_x || !ctx->has_a) return GDSS_E_BADARG;
  if (gdss_device_count() == 0) return GDSS_E_UNSUPPORTED;          // no CPU fallback
  if (scratch_floats < gdss_dsm_scratch_floats(ctx, B)) return GDSS_E_WORKSPACE;
  Workspace w;
  layout_workspace(ctx, B, &w);
  if (workspace_bytes < w.off[R_COUNT]) return GDSS_E_WORKSPACE;
  const int N = ctx->N, F = ctx->F;
  if ((int64_t)B * (N * (N - 1) / 2) >= (int64_t)1 << 30 || B >= (1 << 19)) return GDSS_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float* flags = scratch;
  float* px = flags + (int64_t)B * N;
  float* zx = px + (int64_t)B * N * F;
  float* netx = zx + (int64_t)B * N * F;
  float* padj = netx + (int64_t)B * N * F;
  float* zadj = padj + (int64_t)B * N * N;
  float* neta = zadj + (int64_t)B * N * N;
  float* lossb = neta + (int64_t)B * N * N;
  dsm_perturb_kernel<<<B, 256, sizeof(float) * N, st>>>(x, adj, coef, raw_x, raw_adj, flags, px, padj, zx, zadj, N, F);
  launched(K_MISC, st);
  LAUNCH_CHECK();
  int rc = launch_setup(ctx, flags, B, 1, ws, w, st);
  if (rc) return rc;
  const int* neff = reinterpret_cast<const int*>(ws + w.off[R_NEFF]);
  rc = launch_eval(ctx, px, padj, flags, neff, netx, neta, B, ws, w, st);
  if (rc) return rc;
  dsm_loss_kernel<<<B, 256, 0, st>>>(netx, neta, zx, zadj, coef, lossb, N, F, reduce_mean);
  launched(K_MISC, st);
  LAUNCH_CHECK();
  dsm_mean_kernel<<<1, 1024, 0, st>>>(lossb, B, loss_out);
  launched(K_MISC, st);
  LAUNCH_CHECK();
  return 0;
}
