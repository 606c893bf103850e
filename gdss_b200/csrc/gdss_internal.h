// Internal (non-ABI) definitions shared by the translation units of libgdss_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/gdss_b200.h"

#define GDSS_MAX_LAYERS 16
#define GDSS_MAX_LIN 4
#define GDSS_MAX_STEPS 8192
#define GDSS_MAX_CH 8        // c_init / c_hid / c_final upper bound the edge kernels are built for

namespace gdss {

// One AttentionLayer (models/attention.py:75-116) in the packed blob.  All Linear weights
// are stored transposed, [in][out] row-major, so that consecutive threads (consecutive
// output units) read consecutive floats.
struct LayerPlan {
  int c_in, c_out, f_in, attn, nhid, hid, nlin, qkvw;
  int f_in_p;                   // f_in rounded up to 4 (zero rows): row count of every wqkv channel block
  int64_t wqkv;                 // [c_in][f_in_p][qkvw]   columns = Q | K | V
  int64_t bqkv;                 // [c_in][qkvw]
  int64_t ew[GDSS_MAX_LIN];     // edge MLP `mlp`: (2 c_in -> hid), (hid -> hid)*, (hid -> c_out)
  int64_t eb[GDSS_MAX_LIN];
  int64_t nw[2], nb[2];         // node MLP `multi_channel`: (c_in*nhid -> hid), (hid -> nhid)
  // Dead sub-network elimination (analyze_liveness, gdss_pack.cu): bit c set = the Q/K chain + attention map (qk_live)
  // or the V convolution (v_live) of channel c can reach an output.  plan_net sets everything live.
  uint32_t qk_live, v_live;
  int node_live;                // the layer's node output x_out is read by something downstream
  int edge_const;               // every input weight of `mlp` is dead: its output is the constant edge_cst per pixel
  float edge_cst[GDSS_MAX_CH];
};

struct NetPlan {
  int type;                     // GDSS_NET_*
  int F, N, nhid, depth, nlin, c_init, c_hid, c_final, adim, heads;
  int fdim;                     // input width of `final`
  int fout;                     // output width of `final` (F for X nets, 1 for A)
  int fdp, Hp, foutp;           // fdim, 2 fdim, fout rounded up to 4: padded (zero) rows / row strides of `final`
  int n_tensors;                // number of state_dict tensors
  LayerPlan layers[GDSS_MAX_LAYERS];
  int64_t gw[GDSS_MAX_LAYERS], gb[GDSS_MAX_LAYERS];   // ScoreNetworkX: DenseGCNConv weight [in rounded up to 4][out], bias
  int64_t fw[3], fb[3];         // final MLP (3 linears): [fdp][Hp], [Hp][Hp], [Hp][foutp]; biases [Hp],[Hp],[foutp]
  int64_t total;                // floats
};

int plan_net(const gdss_net_desc* d, int N, NetPlan* p);
// Marks the parts of an AttentionLayer stack that cannot influence the output because whole weight blocks are below
// GDSS_DEAD_EPS in magnitude (weight-decayed dead sub-networks of the shipped checkpoints).  host_blob = the packed blob.
// Returns the number of (layer, channel) chains / MLPs it switched off.
int analyze_liveness(NetPlan* p, const float* host_blob);

// named workspace regions (byte offsets, 256-aligned)
enum Region {
  R_NEFF = 0, R_STEP, R_TABLE, R_MEANS, R_NORMS, R_STACK_A, R_STACK_X, R_DIS, R_QKV, R_XA0, R_XA1, R_XS, R_RAW_X, R_RAW_ADJ,
  R_POFF, R_META, R_GMAP, R_ORDER, R_STACKT, R_X0, R_A0, R_NOFF, R_NMAP,
  R_COUNT
};

struct Workspace {
  int64_t off[R_COUNT + 1];
  int64_t pstride;              // fused path: floats per channel plane of the pixel-major stack
  int64_t nstride;              // fused path: floats per feature plane of the X network's node stack (rows b * N + i)
};

}  // namespace gdss

struct gdss_ctx {
  int N, F;
  bool has_x, has_a;
  gdss::NetPlan px, pa;
  const float* blob_x;
  const float* blob_a;
  int smem_optin;               // max dynamic shared memory per block
  int num_sms;
  bool fused;                   // small-graph path: one CTA per graph (gdss_fused.cu)
  // side streams / events of the size-bucket launches of the per-graph kernel (they run as parallel branches; one call per
  // context in flight at a time)
  cudaStream_t side[2];
  cudaEvent_t ev_fork, ev_join[2];
  bool use_mma;                 // per-graph contractions on mma.sync (3xTF32); GDSS_NO_MMA=1 selects the FFMA variants
  bool use_tc_edge;             // per-edge layer MLPs of the per-graph kernel on tcgen05 (GDSS_NO_TCEDGE=1: the mma.sync version)
  bool use_tc;                  // final edge MLP on tcgen05 tensor cores (3xTF32); GDSS_NO_TC=1 selects the FFMA kernel
  int pruned;                   // number of dead chains / MLPs switched off by analyze_liveness (GDSS_NO_PRUNE=1: 0)
  bool phase_prof;              // GDSS_PHASE_PROF=1 (read once at creation): per-phase cycle counters in the workspace
};

namespace gdss {

void layout_workspace(const gdss_ctx* ctx, int B, Workspace* w);

// Enqueue one joint evaluation of the bound networks (raw outputs).  neff = device int[B].
// zero_fill: clear the outputs first (entries of padded nodes are never written).  The fused path writes valid entries
// only, so a sampler clears its score buffers once and passes false for every step.
int launch_eval(const gdss_ctx* ctx, const float* x, const float* adj, const float* flags, const int* neff,
                float* score_x, float* score_adj, int B, char* ws, const Workspace& w, cudaStream_t st, bool zero_fill = true);
int launch_neff(const float* flags, int* neff, int B, int N, int inputs_masked, cudaStream_t st);

void count_launch(int n = 1);

// optional per-kernel-class timing (gdss_prof_begin / gdss_prof_end): when enabled every launch site
// records a CUDA event right after its launch on the launching stream.
enum KernelId { K_NEFF = 0, K_POW, K_DEG, K_GCN, K_EDGE, K_NODE, K_FINAL_EDGE, K_FINAL_NODE, K_COPY, K_NORMS, K_SUMS,
                K_UPDATE, K_MISC, K_MEMOPS, K_FUSED, K_COUNT };
void prof_tick(int kid, cudaStream_t st);
bool prof_active();
// raise a kernel's dynamic shared-memory limit if (and only if) `bytes` exceeds what it was granted before on this device
int ensure_smem(const void* fn, size_t bytes);
inline void launched(int kid, cudaStream_t st) { count_launch(); prof_tick(kid, st); }

// fused small-graph path (gdss_fused.cu)
bool fused_supported(const gdss_ctx* ctx);
int64_t fused_query(const gdss_ctx* ctx, int what);   // 0: networks fusable, 1: shared-memory bytes at N
int launch_tri_setup(const gdss_ctx* ctx, const int* neff, int B, char* ws, const Workspace& w, cudaStream_t st);
int launch_eval_fused(const gdss_ctx* ctx, const float* x, const float* adj, const float* flags, const int* neff,
                      float* score_x, float* score_adj, int B, char* ws, const Workspace& w, cudaStream_t st);
// general path: ScoreNetworkA.final on the tensor-core kernel over the dense channel stack (needs the pixel bookkeeping)
bool final_tc_dense_ok(const gdss_ctx* ctx, int B);
int launch_final_tc_dense(const gdss_ctx* ctx, const float* stack, int64_t s_bs, const float* flags, float* score_adj, int B,
                          char* ws, const Workspace& w, cudaStream_t st);
// flags -> neff (+ the pixel bookkeeping of the fused path); call once per flags tensor
int launch_setup(const gdss_ctx* ctx, const float* flags, int B, int inputs_masked, char* ws, const Workspace& w,
                 cudaStream_t st);

// losses.py:47-58 (gdss_loss.cu): node flags of the data, masked noise from the raw draws, perturbed inputs
int launch_dsm_perturb(const float* x, const float* adj, const float* coef, const float* raw_x, const float* raw_adj, float* flags,
                       float* px, float* padj, float* zx, float* zadj, int B, int N, int F, cudaStream_t st);

// NCCL (dlopen'ed)
int nccl_allreduce_sum(void* comm, float* buf, int count, cudaStream_t st);

}  // namespace gdss
