// Optimizer half of the reference's training step (trainer.py:57-79) on flat fp32 arrays of one network:
//   torch.nn.utils.clip_grad_norm_(params, grad_norm)   total L2 norm over all gradients, coef = min(1, max / (norm + 1e-6))
//   torch.optim.Adam(lr, weight_decay).step()           L2 weight decay folded into the gradient, bias-corrected moments
//   ExponentialMovingAverage.update (utils/ema.py:18-35) shadow -= (1 - decay) * (shadow - param)
// fused into two launches: a deterministic two-stage reduction of the squared gradient norm, then one elementwise pass over
// (param, grad, exp_avg, exp_avg_sq, shadow).  The data-parallel trainer all-reduces the flat gradient buffer first
// (gdss_nccl_allreduce) so that the clip sees the GLOBAL gradient.  HBM-bound: 5 reads + 4 writes per parameter.
#include "gdss_common.cuh"
#include "gdss_internal.h"

namespace gdss {

#define LAUNCH_CHECK()                          \
  do {                                          \
    cudaError_t e__ = cudaGetLastError();       \
    if (e__ != cudaSuccess) return (int)e__;    \
  } while (0)

#define OPT_BLOCKS 256          // partial sums of the norm reduction (fixed, so the summation order never depends on n)

// stage 1: partial[k] = sum of squares of the k-th contiguous slice (fixed-order tree inside the block)
__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const float* __restrict__ g, int64_t n, float scale,
                                                          float* __restrict__ partial) {
  __shared__ float red[256];
  const int64_t per = (n + OPT_BLOCKS - 1) / OPT_BLOCKS;
  const int64_t lo = blockIdx.x * per, hi = lo + per < n ? lo + per : n;
  float s = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
    const float v = g[i] * scale;
    s = fmaf(v, v, s);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

struct AdamArgs {
  float* p; const float* g; float* m; float* v; float* shadow;
  int64_t n;
  float lr, beta1, beta2, eps, weight_decay, grad_clip, grad_scale, bc1, bc2_sqrt, ema_one_minus_decay;
  const unsigned char* active; // optional per-parameter mask: 0 = the loss never reaches it (grad None in torch: Adam skips it)
  const float* consts;         // optional device float[4] {lr, bc1, bc2_sqrt, ema_one_minus_decay}: overrides the by-value copies
                               // (a captured CUDA graph replays the launch; the host refreshes the four floats per step)
  const float* partial;        // OPT_BLOCKS partial squared norms (grad_clip > 0)
  float* norm_out;             // optional: the total gradient norm (what clip_grad_norm_ returns)
};

// stage 2: every block re-adds the OPT_BLOCKS partials in the same order (bit-identical coefficient in all blocks)
__global__ void __launch_bounds__(256) adam_ema_kernel(AdamArgs a) {
  __shared__ float coef_s;
  if (threadIdx.x == 0) {
    float coef = 1.f;
    if (a.grad_clip > 0.f) {
      float s = 0.f;
      for (int k = 0; k < OPT_BLOCKS; ++k) s += a.partial[k];
      const float norm = sqrtf(s);
      coef = fminf(a.grad_clip / (norm + 1e-6f), 1.f);            // clip_grad_norm_: clamp(max_norm / (total_norm + 1e-6), max=1)
      if (a.norm_out && blockIdx.x == 0) *a.norm_out = norm;
    }
    coef_s = coef * a.grad_scale;
  }
  __syncthreads();
  const float coef = coef_s;
  float lr = a.lr, bc1 = a.bc1, bc2_sqrt = a.bc2_sqrt, omd = a.ema_one_minus_decay;
  if (a.consts) { lr = a.consts[0]; bc1 = a.consts[1]; bc2_sqrt = a.consts[2]; omd = a.consts[3]; }
  const float step_size = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * 256) {
    float p = a.p[i];
    if (a.active && !a.active[i]) {                                // torch.optim.Adam skips parameters without a gradient:
      if (a.shadow) {                                              // no weight decay, no moment update; the EMA still runs
        const float s = a.shadow[i];
        a.shadow[i] = s - omd * (s - p);
      }
      continue;
    }
    float g = a.g[i] * coef;
    if (a.weight_decay != 0.f) g = fmaf(a.weight_decay, p, g);      // Adam (not AdamW): grad += weight_decay * param
    const float m = a.beta1 * a.m[i] + (1.f - a.beta1) * g;        // exp_avg.lerp_(grad, 1 - beta1)
    const float v = a.beta2 * a.v[i] + (1.f - a.beta2) * g * g;    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    a.m[i] = m;
    a.v[i] = v;
    const float denom = sqrtf(v) / bc2_sqrt + a.eps;
    p = p - step_size * (m / denom);
    a.p[i] = p;
    if (a.shadow) {
      const float s = a.shadow[i];
      a.shadow[i] = s - omd * (s - p);
    }
  }
}

}  // namespace gdss

using namespace gdss;

extern "C" int64_t gdss_optim_scratch_floats(void) { return OPT_BLOCKS + 4; }

// {lr, 1 - beta1^step, sqrt(1 - beta2^step), 1 - ema decay of this update}: in double on the host, as torch computes them from the
// python step count (utils/ema.py:26-28: num_updates += 1; decay = min(decay, (1 + num_updates) / (10 + num_updates)))
extern "C" int gdss_adam_step_consts(float lr, float beta1, float beta2, int32_t step, float ema_decay, int64_t ema_num_updates,
                                     float* out4) {
  if (!out4 || step < 1) return GDSS_E_BADARG;
  out4[0] = lr;
  out4[1] = (float)(1.0 - pow((double)beta1, (double)step));
  out4[2] = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  double decay = ema_decay;
  if (ema_num_updates >= 0) {
    const double nu = (double)(ema_num_updates + 1);
    const double warm = (1.0 + nu) / (10.0 + nu);
    if (warm < decay) decay = warm;
  }
  out4[3] = (float)(1.0 - decay);
  return 0;
}

static int adam_launch(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema_shadow, int64_t n,
                       float beta1, float beta2, float eps, float weight_decay, float grad_clip, float grad_scale,
                       const float host4[4], const float* consts_dev, const unsigned char* active, float* scratch,
                       float* grad_norm_out, void* stream) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return GDSS_E_UNSUPPORTED;   // no CPU fallback
  cudaStream_t st = (cudaStream_t)stream;
  if (grad_clip > 0.f) {
    grad_sqnorm_kernel<<<OPT_BLOCKS, 256, 0, st>>>(grads, n, grad_scale, scratch);
    launched(K_MISC, st);
    LAUNCH_CHECK();
  }
  AdamArgs a;
  a.p = params; a.g = grads; a.m = exp_avg; a.v = exp_avg_sq; a.shadow = ema_shadow; a.n = n;
  a.lr = host4[0]; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.grad_clip = grad_clip;
  a.grad_scale = grad_scale;
  a.bc1 = host4[1];
  a.bc2_sqrt = host4[2];
  a.ema_one_minus_decay = host4[3];
  a.consts = consts_dev;
  a.active = active;
  a.partial = scratch;
  a.norm_out = grad_norm_out;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_ema_kernel<<<blocks, 256, 0, st>>>(a);
  launched(K_MISC, st);
  LAUNCH_CHECK();
  return 0;
}

extern "C" int gdss_adam_ema_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema_shadow,
                                  int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, float grad_clip,
                                  float grad_scale, int32_t step, float ema_decay, int64_t ema_num_updates,
                                  const unsigned char* active, float* scratch, float* grad_norm_out, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || n <= 0 || step < 1 || !scratch) return GDSS_E_BADARG;
  float h4[4];
  gdss_adam_step_consts(lr, beta1, beta2, step, ema_decay, ema_num_updates, h4);
  return adam_launch(params, grads, exp_avg, exp_avg_sq, ema_shadow, n, beta1, beta2, eps, weight_decay, grad_clip, grad_scale, h4,
                     nullptr, active, scratch, grad_norm_out, stream);
}

// The same step with the four per-step scalars read from device memory (gdss_adam_step_consts, copied by the caller before the
// launch): the launch itself no longer depends on the step count, so it can sit inside a captured CUDA graph.
extern "C" int gdss_adam_ema_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema_shadow,
                                      int64_t n, float beta1, float beta2, float eps, float weight_decay, float grad_clip,
                                      float grad_scale, const float* step_consts_dev, const unsigned char* active, float* scratch,
                                      float* grad_norm_out, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || n <= 0 || !step_consts_dev || !scratch) return GDSS_E_BADARG;
  const float h4[4] = {0.f, 1.f, 1.f, 0.f};
  return adam_launch(params, grads, exp_avg, exp_avg_sq, ema_shadow, n, beta1, beta2, eps, weight_decay, grad_clip, grad_scale, h4,
                     step_consts_dev, active, scratch, grad_norm_out, stream);
}
