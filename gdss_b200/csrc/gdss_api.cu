// extern "C" entry points that are not the solver: library info, context, score networks,
// tensor helpers (utils/graph_utils.py), NCCL plumbing.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gdss_common.cuh"
#include "gdss_internal.h"

using namespace gdss;

#define LAUNCH_CHECK()                          \
  do {                                          \
    cudaError_t e__ = cudaGetLastError();       \
    if (e__ != cudaSuccess) return (int)e__;    \
  } while (0)

extern "C" int gdss_abi_version(void) { return GDSS_ABI_VERSION; }

extern "C" int gdss_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" const char* gdss_error_string(int code) {
  switch (code) {
    case GDSS_OK: return "ok";
    case GDSS_E_BADARG: return "bad argument (null pointer, negative size or shape mismatch)";
    case GDSS_E_UNSUPPORTED: return "configuration not supported by the sm_100a kernels (or no CUDA device: there is no CPU fallback)";
    case GDSS_E_WORKSPACE: return "workspace too small";
    case GDSS_E_NCCL: return "NCCL not loadable or collective failed";
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown error";
}

extern "C" gdss_ctx* gdss_ctx_create(int32_t N, int32_t F, const gdss_net_desc* net_x, const float* dev_blob_x,
                                     const gdss_net_desc* net_a, const float* dev_blob_a) {
  if (N <= 0 || F <= 0 || N >= 65536) return nullptr;
  gdss_ctx* c = (gdss_ctx*)calloc(1, sizeof(gdss_ctx));
  if (!c) return nullptr;
  c->N = N;
  c->F = F;
  if (net_x) {
    if (net_x->model_type == GDSS_NET_A || net_x->max_feat_num != F || plan_net(net_x, N, &c->px) < 0 || !dev_blob_x) {
      free(c);
      return nullptr;
    }
    c->has_x = true;
    c->blob_x = dev_blob_x;
  }
  if (net_a) {
    if (net_a->model_type != GDSS_NET_A || net_a->max_feat_num != F || plan_net(net_a, N, &c->pa) < 0 || !dev_blob_a) {
      free(c);
      return nullptr;
    }
    c->has_a = true;
    c->blob_a = dev_blob_a;
  }
  c->smem_optin = 48 * 1024;
  c->num_sms = 0;
  int dev = 0;
  if (gdss_device_count() > 0 && cudaGetDevice(&dev) == cudaSuccess) {
    cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (c->num_sms > 0 && !(getenv("GDSS_NO_FORK") && getenv("GDSS_NO_FORK")[0] == '1')) {
    bool ok = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    // The side streams carry the LARGER size buckets of the per-graph kernel at a higher priority than the caller's stream
    // (which carries the smallest graphs): the block scheduler then places the big graphs first and the launch ends with the
    // smallest ones -- a short last wave (GDSS_NO_PRIO=1: equal priorities)
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);          // numerically lower = higher priority
    const bool noprio = getenv("GDSS_NO_PRIO") && getenv("GDSS_NO_PRIO")[0] == '1';
    for (int k = 0; k < 2 && ok; ++k) {
      int pr = prio_hi + k;
      if (pr > prio_lo || noprio) pr = prio_lo;
      ok = cudaStreamCreateWithPriority(&c->side[k], cudaStreamNonBlocking, pr) == cudaSuccess &&
           cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) c->side[0] = c->side[1] = nullptr;
  }
  c->fused = false;
  const char* nf = getenv("GDSS_NO_FUSED");
  if (!(nf && nf[0] == '1')) c->fused = fused_supported(c);
  const char* ntc = getenv("GDSS_NO_TC");
  c->use_tc = !(ntc && ntc[0] == '1');
  const char* nmma = getenv("GDSS_NO_MMA");
  c->use_mma = !(nmma && nmma[0] == '1');
  const char* te = getenv("GDSS_TCEDGE");          // opt-in experiment: per-edge layer MLPs on tcgen05 (slower)
  c->use_tc_edge = te && te[0] == '1';
  const char* pp = getenv("GDSS_PHASE_PROF");
  c->phase_prof = pp && pp[0] == '1';
  // dead sub-network elimination (gdss_pack.cu analyze_liveness): needs the weights on the host once
  const char* np = getenv("GDSS_NO_PRUNE");
  if (c->num_sms > 0 && !(np && np[0] == '1')) {
    const float* blobs[2] = {c->has_x ? dev_blob_x : nullptr, c->has_a ? dev_blob_a : nullptr};
    NetPlan* plans[2] = {&c->px, &c->pa};
    for (int k = 0; k < 2; ++k) {
      if (!blobs[k] || plans[k]->type == GDSS_NET_X) continue;
      float* host = (float*)malloc(sizeof(float) * (size_t)plans[k]->total);
      if (!host) continue;
      if (cudaMemcpy(host, blobs[k], sizeof(float) * (size_t)plans[k]->total, cudaMemcpyDeviceToHost) == cudaSuccess)
        c->pruned += analyze_liveness(plans[k], host);
      else
        cudaGetLastError();
      free(host);
    }
  }
  return c;
}

extern "C" void gdss_ctx_destroy(gdss_ctx* ctx) {
  if (!ctx) return;
  for (int k = 0; k < 2; ++k) {
    if (ctx->side[k]) cudaStreamDestroy(ctx->side[k]);
    if (ctx->ev_join[k]) cudaEventDestroy(ctx->ev_join[k]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  free(ctx);
}

extern "C" int64_t gdss_workspace_bytes(const gdss_ctx* ctx, int32_t B) {
  if (!ctx || B <= 0) return GDSS_E_BADARG;
  Workspace w;
  layout_workspace(ctx, B, &w);
  return w.off[R_COUNT];
}

extern "C" int64_t gdss_workspace_offset(const gdss_ctx* ctx, int32_t B, const char* region) {
  static const char* names[R_COUNT] = {"neff", "step", "table", "means", "norms", "stack_a", "stack_x", "dis",
                                       "qkv", "xa0", "xa1", "xs", "raw_x", "raw_adj", "poff", "meta", "gmap", "order", "stackT", "x0", "a0", "noff", "nmap"};
  if (!ctx || B <= 0 || !region) return GDSS_E_BADARG;
  Workspace w;
  layout_workspace(ctx, B, &w);
  for (int i = 0; i < R_COUNT; ++i)
    if (!strcmp(names[i], region)) return w.off[i];
  return GDSS_E_BADARG;
}

extern "C" int64_t gdss_ctx_query(const gdss_ctx* ctx, const char* key) {
  if (!ctx || !key) return GDSS_E_BADARG;
  if (!strcmp(key, "fused")) return ctx->fused ? 1 : 0;
  if (!strcmp(key, "fused_fusable")) return fused_query(ctx, 0);
  if (!strcmp(key, "fused_smem_bytes")) return fused_query(ctx, 1);
  if (!strncmp(key, "bucket_smem_bytes:", 18)) {          // "bucket_smem_bytes:<NB>:<0|1>" (nodes of the bucket, weights staged)
    int nb = 0, wsm = 0;
    if (sscanf(key + 18, "%d:%d", &nb, &wsm) != 2) return GDSS_E_BADARG;
    return fused_query(ctx, 1000 + 2 * nb + (wsm ? 1 : 0));
  }
  if (!strcmp(key, "smem_optin")) return ctx->smem_optin;
  if (!strcmp(key, "use_tc")) return ctx->use_tc ? 1 : 0;
  if (!strcmp(key, "use_mma")) return ctx->use_mma ? 1 : 0;
  if (!strcmp(key, "pruned")) return ctx->pruned;
  return GDSS_E_BADARG;
}

extern "C" int gdss_score_forward(gdss_ctx* ctx, const float* x, const float* adj, const float* flags, float* score_x,
                                  float* score_adj, int32_t B, int32_t inputs_masked, void* workspace,
                                  int64_t workspace_bytes, void* stream) {
  if (!ctx || !x || !adj || !flags || !workspace || B <= 0) return GDSS_E_BADARG;
  if ((score_x && !ctx->has_x) || (score_adj && !ctx->has_a)) return GDSS_E_BADARG;
  if (gdss_device_count() == 0) return GDSS_E_UNSUPPORTED;     // no CPU fallback
  Workspace w;
  layout_workspace(ctx, B, &w);
  if (workspace_bytes < w.off[R_COUNT]) return GDSS_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  int* neff = reinterpret_cast<int*>(ws + w.off[R_NEFF]);
  if ((int64_t)B * (ctx->N * (ctx->N - 1) / 2) >= (int64_t)1 << 30 || B >= (1 << 19)) return GDSS_E_UNSUPPORTED;
  int rc = launch_setup(ctx, flags, B, inputs_masked, ws, w, st);
  if (rc) return rc;
  return launch_eval(ctx, x, adj, flags, neff, score_x, score_adj, B, ws, w, st);
}

// ---------------------------------------------------------------------------------------
// tensor helpers (utils/graph_utils.py)
// ---------------------------------------------------------------------------------------
namespace {

__global__ void mask_x_kernel(float* x, const float* flags, int64_t rows, int F) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < rows * F; idx += (int64_t)gridDim.x * blockDim.x)
    x[idx] *= flags[idx / F];
}

// adjs [B][C][N][N] *= flags[b][i] then *= flags[b][j]   (graph_utils.py:16-29)
__global__ void mask_adjs_kernel(float* a, const float* flags, int64_t B, int C, int N) {
  const int64_t total = B * C * N * N;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int j = idx % N;
    int i = (idx / N) % N;
    int64_t b = idx / ((int64_t)C * N * N);
    a[idx] = a[idx] * flags[b * N + i] * flags[b * N + j];
  }
}

// gen_noise (graph_utils.py:78-86) from a supplied raw draw or the library's Philox stream
__global__ void gen_noise_kernel(float* out, const float* raw, const float* flags, int B, int N, int F, int sym,
                                 uint64_t seed, uint32_t draw, int64_t graph_offset) {
  const int b = blockIdx.x;
  const float* fl = flags + (int64_t)b * N;
  if (!sym) {
    const int gw = (F + 3) >> 2;
    for (int idx = threadIdx.x; idx < N * gw; idx += blockDim.x) {
      int i = idx / gw, g = idx - i * gw;
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!raw) z = philox_normal4(seed, (uint64_t)(graph_offset + b), draw, 0u, (uint32_t)(i * gw + g));
      for (int q = 0; q < 4; ++q) {
        int f = 4 * g + q;
        if (f >= F) continue;
        int64_t e = ((int64_t)b * N + i) * F + f;
        float v = raw ? raw[e] : f4_get(z, q);
        out[e] = v * fl[i];
      }
    }
  } else {
    const int gw = (N + 3) >> 2;
    for (int i = threadIdx.x; i < N; i += blockDim.x) out[((int64_t)b * N + i) * N + i] = 0.f;
    for (int idx = threadIdx.x; idx < N * gw; idx += blockDim.x) {
      int i = idx / gw, g = idx - i * gw;
      if (4 * g + 3 <= i) continue;
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!raw) z = philox_normal4(seed, (uint64_t)(graph_offset + b), draw, 1u, (uint32_t)(i * gw + g));
      for (int q = 0; q < 4; ++q) {
        int j = 4 * g + q;
        if (j <= i || j >= N) continue;
        float v = raw ? raw[((int64_t)b * N + i) * N + j] : f4_get(z, q);
        v = v * fl[i] * fl[j];
        out[((int64_t)b * N + i) * N + j] = v;
        out[((int64_t)b * N + j) * N + i] = v;
      }
    }
  }
}

__global__ void quantize_kernel(const float* a, float* out, float thr, int64_t n) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x)
    out[idx] = a[idx] < thr ? 0.f : 1.f;       // torch.where(adjs < thr, 0, 1): NaN -> 1, like the reference
}

__global__ void quantize_mol_kernel(const float* a, int64_t* out, int64_t n) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    // graph_utils.py:101-104: three successive torch.where passes; a NaN fails every comparison and the reference
    // then casts the NaN itself to int64 (undefined); here NaN -> 0 (documented, asserted in the host mirror)
    float v = a[idx];
    out[idx] = v >= 2.5f ? 3 : (v >= 1.5f ? 2 : (v >= 0.5f ? 1 : 0));
  }
}

// Post-sampling decode of a molecule batch (sampler.py:131-138) in one pass, compact int8 outputs:
//   bond[b][i][j]  = class index of quantize_mol(adj) - 1 with -1 -> 3  (0 single, 1 double, 2 triple, 3 no bond)
//   atom[b][i][f]  = x > 0.5 for f < F;  atom[b][i][F] = 1 - sum_f atom[b][i][f]   (the virtual-atom channel)
__global__ void decode_mol_kernel(const float* __restrict__ x, const float* __restrict__ adj, int8_t* __restrict__ atom,
                                  int8_t* __restrict__ bond, int64_t rows, int N, int F) {
  const int64_t nb = rows * N;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < nb; idx += (int64_t)gridDim.x * blockDim.x) {
    const float v = adj[idx];
    const int q = v >= 2.5f ? 3 : (v >= 1.5f ? 2 : (v >= 0.5f ? 1 : 0));
    bond[idx] = (int8_t)(q == 0 ? 3 : q - 1);
  }
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    int s = 0;
    for (int f = 0; f < F; ++f) {
      const int a = x[r * F + f] > 0.5f ? 1 : 0;
      atom[r * (F + 1) + f] = (int8_t)a;
      s += a;
    }
    atom[r * (F + 1) + F] = (int8_t)(1 - s);
  }
}

// init_flags (utils/graph_utils.py:66-74) on the device: the reference draws a training graph uniformly and takes its
// node flags, i.e. a node count from the training set's histogram.  cdf = inclusive cumulative counts of node numbers
// 0..N (cdf[N] = number of training graphs); graph b draws u from the Philox stream (domain 2) and gets the first
// n_b flags set, n_b = min { k : u * cdf[N] < cdf[k] }.
__global__ void init_flags_kernel(const float* __restrict__ cdf, float* __restrict__ flags, int B, int N, uint64_t seed,
                                  int64_t graph_offset) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint64_t gid = (uint64_t)(graph_offset + b);
  uint4 c = make_uint4(0u, (uint32_t)gid, ((uint32_t)(gid >> 32) & 0x3fffffffu) | (2u << 30), 0u);
  const uint4 r = philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float u = (float)(r.x >> 8) * (1.0f / 16777216.0f) * cdf[N];        // [0, total)
  int n = N;
  for (int k = 0; k <= N; ++k)
    if (u < cdf[k]) { n = k; break; }
  for (int i = 0; i < N; ++i) flags[(int64_t)b * N + i] = i < n ? 1.f : 0.f;
}

// full (unmirrored) batched product for the standalone pow_tensor helper
__global__ void __launch_bounds__(256) bmm_kernel(const float* __restrict__ Lp, int64_t l_bs, const float* __restrict__ Ap,
                                                  int64_t a_bs, float* __restrict__ out, int64_t o_bs, int N) {
  const int b = blockIdx.z, ti = blockIdx.y, tj = blockIdx.x;
  __shared__ float Ls[32][33], Rs[32][33];
  const float* L = Lp + b * l_bs;
  const float* A = Ap + b * a_bs;
  float* O = out + b * o_bs;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < N; k0 += 32) {
    for (int q = 0; q < 4; ++q) {
      int r = ty + 8 * q;
      int gi = ti * 32 + r, gk = k0 + tx;
      Ls[r][tx] = (gi < N && gk < N) ? L[(int64_t)gi * N + gk] : 0.f;
      int gr = k0 + r, gj = tj * 32 + tx;
      Rs[r][tx] = (gr < N && gj < N) ? A[(int64_t)gr * N + gj] : 0.f;
    }
    __syncthreads();
    for (int k = 0; k < 32; ++k) {
      float rv = Rs[k][tx];
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(Ls[ty + 8 * q][k], rv, acc[q]);
    }
    __syncthreads();
  }
  for (int q = 0; q < 4; ++q) {
    int i = ti * 32 + ty + 8 * q, j = tj * 32 + tx;
    if (i < N && j < N) O[(int64_t)i * N + j] = acc[q];
  }
}

inline int grid_for(int64_t n) {
  int64_t g = (n + 255) / 256;
  return (int)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g));
}

}  // namespace

extern "C" int gdss_mask_x(float* x, const float* flags, int32_t B, int32_t N, int32_t F, void* stream) {
  if (!x || !flags || B <= 0 || N <= 0 || F <= 0) return GDSS_E_BADARG;
  if (gdss_device_count() == 0) return GDSS_E_UNSUPPORTED;
  mask_x_kernel<<<grid_for((int64_t)B * N * F), 256, 0, (cudaStream_t)stream>>>(x, flags, (int64_t)B * N, F);
  count_launch();
  LAUNCH_CHECK();
  return 0;
}

extern "C" int gdss_mask_adjs(float* adjs, const float* flags, int32_t B, int32_t C, int32_t N, void* stream) {
  if (!adjs || !flags || B <= 0 || N <= 0 || C <= 0) return GDSS_E_BADARG;
  if (gdss_device_count() == 0) return GDSS_E_UNSUPPORTED;
  mask_adjs_kernel<<<grid_for((int64_t)B * C * N * N), 256, 0, (cudaStream_t)stream>>>(adjs, flags, B, C, N);
  count_launch();
  LAUNCH_CHECK();
  return 0;
}

extern "C" int gdss_gen_noise(float* out, const float* raw, const float* flags, int32_t B, int32_t N, int32_t F,
                              int32_t sym, uint64_t seed, uint64_t draw_id, int64_t graph_offset, void* stream) {
  if (!out || !flags || B <= 0 || N <= 0 || (!sym && F <= 0)) return GDSS_E_BADARG;
  if (gdss_device_count() == 0) return GDSS_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (sym) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * N * N, st);
    if (e != cudaSuccess) return (int)e;
  }
  gen_noise_kernel<<<B, 256, 0, st>>>(out, raw, flags, B, N, F, sym, seed, (uint32_t)draw_id, graph_offset);
  count_launch();
  LAUNCH_CHECK();
  return 0;
}

extern "C" int gdss_pow_tensor(const float* adj, float* out, int32_t B, int32_t N, int32_t cnum, void* stream) {
  if (!adj || !out || B <= 0 || N <= 0 || cnum <= 0) return GDSS_E_BADARG;
  if (gdss_device_count() == 0) return GDSS_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t NN = (int64_t)N * N;
  cudaError_t e = cudaMemcpy2DAsync(out, cnum * NN * 4, adj, NN * 4, NN * 4, B, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return (int)e;
  const int nt = (N + 31) / 32;
  for (int k = 1; k < cnum; ++k) {
    bmm_kernel<<<dim3(nt, nt, B), 256, 0, st>>>(out + (k - 1) * NN, cnum * NN, adj, NN, out + k * NN, cnum * NN, N);
    count_launch();
    LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int gdss_quantize(const float* adj, float* out, float thr, int64_t numel, void* stream) {
  if (!adj || !out || numel <= 0) return GDSS_E_BADARG;
  if (gdss_device_count() == 0) return GDSS_E_UNSUPPORTED;
  quantize_kernel<<<grid_for(numel), 256, 0, (cudaStream_t)stream>>>(adj, out, thr, numel);
  count_launch();
  LAUNCH_CHECK();
  return 0;
}

extern "C" int gdss_quantize_mol(const float* adj, int64_t* out, int64_t numel, void* stream) {
  if (!adj || !out || numel <= 0) return GDSS_E_BADARG;
  if (gdss_device_count() == 0) return GDSS_E_UNSUPPORTED;
  quantize_mol_kernel<<<grid_for(numel), 256, 0, (cudaStream_t)stream>>>(adj, out, numel);
  count_launch();
  LAUNCH_CHECK();
  return 0;
}

extern "C" int gdss_decode_mol(const float* x, const float* adj, int8_t* atom, int8_t* bond, int32_t B, int32_t N, int32_t F,
                               void* stream) {
  if (!x || !adj || !atom || !bond || B <= 0 || N <= 0 || F <= 0) return GDSS_E_BADARG;
  if (gdss_device_count() == 0) return GDSS_E_UNSUPPORTED;
  const int64_t rows = (int64_t)B * N;
  decode_mol_kernel<<<grid_for(rows * N), 256, 0, (cudaStream_t)stream>>>(x, adj, atom, bond, rows, N, F);
  count_launch();
  LAUNCH_CHECK();
  return 0;
}

extern "C" int gdss_init_flags(const float* node_count_cdf, float* flags, int32_t B, int32_t N, uint64_t seed,
                               int64_t graph_offset, void* stream) {
  if (!node_count_cdf || !flags || B <= 0 || N <= 0) return GDSS_E_BADARG;
  if (gdss_device_count() == 0) return GDSS_E_UNSUPPORTED;
  init_flags_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(node_count_cdf, flags, B, N, seed, graph_offset);
  count_launch();
  LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------
// NCCL, resolved at run time (the PyTorch host process has libnccl.so.2 loaded already)
// ---------------------------------------------------------------------------------------
namespace {

typedef int (*nccl_get_unique_id_t)(void*);
struct Id128 { char b[128]; };
typedef int (*nccl_init_t)(void**, int, Id128, int);
typedef int (*nccl_destroy_t)(void*);
typedef int (*nccl_allreduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_allgather_t)(const void*, void*, size_t, int, void*, cudaStream_t);

struct Nccl {
  void* h = nullptr;
  nccl_get_unique_id_t get_id = nullptr;
  nccl_init_t init = nullptr;
  nccl_destroy_t destroy = nullptr;
  nccl_allreduce_t allreduce = nullptr;
  nccl_allgather_t allgather = nullptr;
  bool tried = false, ok = false;
};
Nccl g_nccl;

bool load_nccl() {
  if (g_nccl.tried) return g_nccl.ok;
  g_nccl.tried = true;
  const char* names[] = {getenv("GDSS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) return false;
  g_nccl.get_id = (nccl_get_unique_id_t)dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.init = (nccl_init_t)dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.destroy = (nccl_destroy_t)dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.allreduce = (nccl_allreduce_t)dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.allgather = (nccl_allgather_t)dlsym(g_nccl.h, "ncclAllGather");
  g_nccl.ok = g_nccl.get_id && g_nccl.init && g_nccl.destroy && g_nccl.allreduce && g_nccl.allgather;
  return g_nccl.ok;
}

}  // namespace

namespace gdss {
int nccl_allreduce_sum(void* comm, float* buf, int count, cudaStream_t st) {
  if (!load_nccl()) return GDSS_E_NCCL;
  // ncclFloat32 = 7, ncclSum = 0
  return g_nccl.allreduce(buf, buf, (size_t)count, 7, 0, comm, st) == 0 ? 0 : GDSS_E_NCCL;
}
}  // namespace gdss

extern "C" int gdss_nccl_unique_id(void* id128) {
  if (!id128) return GDSS_E_BADARG;
  if (!load_nccl()) return GDSS_E_NCCL;
  return g_nccl.get_id(id128) == 0 ? 0 : GDSS_E_NCCL;
}

extern "C" int gdss_nccl_comm_create(const void* id128, int32_t world, int32_t rank, void** comm_out) {
  if (!id128 || !comm_out || world <= 0 || rank < 0 || rank >= world) return GDSS_E_BADARG;
  if (!load_nccl()) return GDSS_E_NCCL;
  Id128 id;
  memcpy(&id, id128, sizeof(id));
  return g_nccl.init(comm_out, world, id, rank) == 0 ? 0 : GDSS_E_NCCL;
}

extern "C" int gdss_nccl_comm_destroy(void* comm) {
  if (!comm) return GDSS_E_BADARG;
  if (!load_nccl()) return GDSS_E_NCCL;
  return g_nccl.destroy(comm) == 0 ? 0 : GDSS_E_NCCL;
}

extern "C" int gdss_nccl_allreduce(void* comm, float* buf, int64_t count, void* stream) {
  if (!comm || !buf || count <= 0 || count > 0x7fffffff) return GDSS_E_BADARG;
  return nccl_allreduce_sum(comm, buf, (int)count, (cudaStream_t)stream);
}

extern "C" int gdss_nccl_allgather(void* comm, const float* send, float* recv, int64_t count_per_rank, void* stream) {
  if (!comm || !send || !recv || count_per_rank <= 0) return GDSS_E_BADARG;
  if (!load_nccl()) return GDSS_E_NCCL;
  return g_nccl.allgather(send, recv, (size_t)count_per_rank, 7, comm, (cudaStream_t)stream) == 0 ? 0 : GDSS_E_NCCL;
}
