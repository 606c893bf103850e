/*
 * gdss_b200.h -- C ABI of the B200-native GDSS reverse-diffusion sampling hot path.
 *
 * One shared library (gdss_b200/csrc/libgdss_b200.so, CUDA, sm_100a only).  Plain pointers
 * and sizes; no torch / C++ types.  The reference (harryjo97/GDSS) is pure Python/PyTorch
 * and has no FFI of its own: each entry point below names the reference interface it
 * replaces (file:line relative to the upstream repository).  INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every device buffer (weights blob, workspace, state, outputs) is owned by the caller;
 *     the library never allocates device memory and never synchronises the device except
 *     where stated; all work is enqueued on the `stream` argument (a cudaStream_t passed as
 *     void*), so the functions are re-entrant per (ctx, workspace, stream);
 *   - tensors are contiguous row-major fp32 exactly as the reference lays them out:
 *     x [B,N,F], adj [B,N,N], flags [B,N] (0/1 floats);
 *   - return value: 0 = ok, <0 = GDSS_E_* below, >0 = a cudaError_t from the launch;
 *   - there is no CPU fallback: without a CUDA device every compute call returns an error.
 */
#ifndef GDSS_B200_H
#define GDSS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDSS_ABI_VERSION 1

enum {
  GDSS_OK = 0,
  GDSS_E_BADARG = -1,       /* null pointer / negative size / shape mismatch            */
  GDSS_E_UNSUPPORTED = -2,  /* configuration outside what the kernels are built for     */
  GDSS_E_WORKSPACE = -3,    /* workspace too small                                      */
  GDSS_E_NCCL = -4          /* NCCL not loadable / collective failed                    */
};

/* utils/loader.py:37-48 model_type strings */
enum { GDSS_NET_X = 0 /* ScoreNetworkX */, GDSS_NET_X_GMH = 1 /* ScoreNetworkX_GMH */, GDSS_NET_A = 2 /* ScoreNetworkA */ };
/* utils/loader.py:94-108 sde.type strings */
enum { GDSS_SDE_VP = 0, GDSS_SDE_VE = 1, GDSS_SDE_SUBVP = 2 };
/* solver.py:163-164, utils/loader.py:128-131 */
enum { GDSS_PRED_EULER = 0, GDSS_PRED_REVERSE = 1, GDSS_PRED_S4 = 2 };
enum { GDSS_CORR_NONE = 0, GDSS_CORR_LANGEVIN = 1 };

/* Constructor kwargs of the score networks (the `params_x` / `params_adj` dicts stored in
 * every checkpoint, utils/loader.py:150-165).  Unused fields are ignored per model_type. */
typedef struct gdss_net_desc {
  int32_t model_type;    /* GDSS_NET_*                                                    */
  int32_t max_feat_num;  /* F                                                             */
  int32_t max_node_num;  /* N (ScoreNetworkA only; X nets take N from the call)           */
  int32_t nhid;
  int32_t depth;         /* X nets: depth; ScoreNetworkA: num_layers                      */
  int32_t num_linears;
  int32_t c_init, c_hid, c_final;
  int32_t adim;
  int32_t num_heads;
  int32_t reserved;
} gdss_net_desc;

/* Sampler configuration = the arguments of get_pc_sampler / S4_solver (solver.py:153-156,
 * :200-203) plus the two SDEs built by load_sde (utils/loader.py:94-108). */
typedef struct gdss_sde_desc {
  int32_t type;          /* GDSS_SDE_*                                                    */
  int32_t num_scales;    /* N of the SDE (1000 in every shipped config)                   */
  double lo, hi;         /* beta_min,beta_max (VP/subVP) or sigma_min,sigma_max (VE)      */
} gdss_sde_desc;

typedef struct gdss_sampler_desc {
  gdss_sde_desc sde_x, sde_adj;
  int32_t predictor;     /* GDSS_PRED_*                                                   */
  int32_t corrector;     /* GDSS_CORR_* (ignored by S4, which always corrects)            */
  double snr, scale_eps, eps;
  int32_t n_steps;       /* Langevin inner steps of get_pc_sampler (solver.py:126, :138; 1 in every shipped config) */
  int32_t probability_flow;  /* informational: the ODE variant lives in the table (pb halved, pc = 0)   */
  int32_t denoise;       /* return *_mean (noise_removal)                                 */
  int32_t reserved;
} gdss_sampler_desc;

/* One row of the per-step coefficient table (24 floats).  Host-computed once per sampler in the
 * reference's own fp32 expressions (gdss_b200/schedule.py cites the sde.py / solver.py lines) and
 * handed to gdss_sample.  Index 0 = node features x, index 1 = adjacency.  `raw` is the raw
 * network output, the score is s = cs * raw (losses.py:13-29). */
typedef struct gdss_step_coef {
  float cs[2];      /* -1/std(t) (VP, subVP) or 1 (VE)                                      */
  float alpha[2];   /* Langevin alpha_k (solver.py:120-124, :240-243); 1 where not used     */
  float pa[2], pb[2], pc[2];  /* predictor: y_mean = pa*y + pb*raw ; y = y_mean + pc*z       */
  float mu1[2], sig1[2], drift[2], mu2[2], sig2[2];
                    /* S4: y = mu1*y + sig1*z2 ; y += drift*raw ; y_mean = mu2*y ; y = y_mean + sig2*z3 */
  float snr, seps;  /* Langevin signal-to-noise ratio and scale_eps                         */
  float t;          /* time of this step (diagnostic)                                       */
  float reserved;
} gdss_step_coef;

typedef struct gdss_ctx gdss_ctx;   /* opaque: two packed networks + launch plan */

/* ---- library ------------------------------------------------------------------------ */
int gdss_abi_version(void);
/* number of CUDA devices visible (0 on a CPU box); never fails */
int gdss_device_count(void);
const char* gdss_error_string(int code);

/* ---- weights ------------------------------------------------------------------------
 * Replaces: nn.Module construction + load_state_dict (utils/loader.py:37-48, :186-196).
 * `tensors[i]` = i-th tensor of the reference state_dict in state_dict order (host, fp32,
 * contiguous, 'module.' prefix irrelevant); `numel[i]` its element count (validated).
 * gdss_packed_floats returns the blob size in floats (<0 on error).  gdss_pack_weights
 * fills a HOST blob of that size; the caller uploads it to the device unchanged. */
int64_t gdss_packed_floats(const gdss_net_desc* desc, int32_t n_nodes);
int32_t gdss_num_state_tensors(const gdss_net_desc* desc);
int gdss_pack_weights(const gdss_net_desc* desc, int32_t n_nodes, const float* const* tensors,
                      const int64_t* numel, int32_t n_tensors, float* host_blob, int64_t blob_floats);

/* Dead sub-network report of a packed HOST blob (what gdss_ctx_create finds and switches off; gdss_pack.cu
 * analyze_liveness): out[4*l + {0,1,2,3}] = {bit mask of channels whose Q/K chain + attention map is live, bit mask of
 * channels whose V convolution is live, the layer's node output is read downstream, the layer's per-edge MLP output is
 * a constant} for the AttentionLayer l (attention.py:93-116).  Weight blocks with max |w| < 1e-12 are dead (the shipped
 * checkpoints' weight-decayed sub-networks, 1e-14 .. 1e-40: their contribution is below 1/60 of an fp32 ulp).  Returns the number of chains / MLPs switched off (<0 on error). */
int gdss_analyze_weights(const gdss_net_desc* desc, int32_t n_nodes, const float* host_blob, int32_t* out, int32_t max_layers);

/* ---- context ------------------------------------------------------------------------
 * A context binds up to two packed networks (node-feature net, adjacency net; either may
 * be NULL) for graphs of N nodes and F features.  Allocates host memory only; reads the two blobs back once (a
 * synchronous device-to-host copy) for the dead sub-network analysis (GDSS_NO_PRUNE=1 skips it). */
gdss_ctx* gdss_ctx_create(int32_t N, int32_t F, const gdss_net_desc* net_x, const float* dev_blob_x,
                          const gdss_net_desc* net_a, const float* dev_blob_a);
void gdss_ctx_destroy(gdss_ctx* ctx);
/* bytes of device workspace the calls below need for a batch of B graphs */
int64_t gdss_workspace_bytes(const gdss_ctx* ctx, int32_t B);

/* ---- score networks -----------------------------------------------------------------
 * Replaces: ScoreNetworkX.forward / ScoreNetworkX_GMH.forward (models/ScoreNetwork_X.py:32-46,
 * :75-89) and ScoreNetworkA.forward (models/ScoreNetwork_A.py:131-150): raw network outputs
 * (before the -1/std scaling of losses.py:13-21).  score_x / score_adj may be NULL to skip a
 * network.  `inputs_masked` != 0 promises x, adj are already zero on flag==0 rows/cols (true
 * inside the sampler) and lets the kernels skip padded nodes; 0 evaluates all N nodes.
 * adj must be symmetric (it always is in GDSS). */
int gdss_score_forward(gdss_ctx* ctx, const float* x, const float* adj, const float* flags,
                       float* score_x, float* score_adj, int32_t B, int32_t inputs_masked,
                       void* workspace, int64_t workspace_bytes, void* stream);

/* ---- solver ------------------------------------------------------------------------- */
/* gdss_sample: replaces the closure returned by get_pc_sampler / S4_solver
 * (solver.py:158-195, :205-279): runs reverse-diffusion steps first_step .. first_step+n_iters-1
 * (rows of `table`, which holds first_step+n_iters rows at least) on the B graphs of this rank.
 * With use_graph != 0 the step is captured into a CUDA graph and the call synchronises `stream`
 * before it returns; otherwise it only enqueues work.
 *   x, adj        in/out state.  If init_prior != 0 they are overwritten with the masked
 *                 N(0,1) prior (solver.py:174-178) from the Philox stream, else used as given.
 *   x_mean, adj_mean  out: the predictor means of the last step (what denoise=True returns)
 *   inj_x, inj_adj    optional injected raw N(0,1) draws, layouts [n_iters][dx][B][N][F] and
 *                 [n_iters][da][B][N][N] in the reference's draw order (dx=da= n_steps+1 PC+Langevin:
 *                 the corrector iterations' draws then the predictor's, 1 PC without corrector, 3 S4); NULL -> counter-based Philox4x32-10 keyed by
 *                 (seed, global graph id = graph_offset + b, step, draw), so results do not
 *                 depend on how the batch is sharded.
 *   trace         optional [n_iters][8] floats: per step the Langevin step sizes
 *                 (eps_x, eps_adj) and the four batch-mean norms, for parity tests.
 *   global_B, nccl_comm   batch-mean coupling of the Langevin step size (solver.py:130-132):
 *                 with nccl_comm != NULL the four per-step norm sums are all-reduced over
 *                 the communicator and divided by global_B; with NULL the local batch is
 *                 the whole batch (global_B must equal B). */
int gdss_sample(gdss_ctx* ctx, const gdss_sampler_desc* s, const gdss_step_coef* table, int32_t first_step,
                int32_t n_iters, float* x, float* adj, const float* flags, float* x_mean, float* adj_mean, int32_t B,
                int32_t init_prior, uint64_t seed, int64_t graph_offset, int64_t global_B,
                const float* inj_x, const float* inj_adj, float* trace, void* nccl_comm, int32_t use_graph,
                void* workspace, int64_t workspace_bytes, void* stream);

/* gdss_solver_phase: ONE phase of ONE solver step -- the corrector's update_fn pair (LangevinCorrector.update_fn, solver.py:113-149,
 * as pc_sampler calls it at :183-189) or the predictor's (ReverseDiffusionPredictor / EulerMaruyamaPredictor.update_fn,
 * solver.py:45-89, called at :190-193), or both (= what gdss_sample runs per step).  `row` = the coefficient row of step
 * `step` (host); x / adj are updated in place, x_mean / adj_mean are written by the predictor phase (initialised to the state);
 * the noise draws are those of gdss_sample for that step (Philox keyed by seed / graph / step / draw, or inj_x / inj_adj =
 * [draws][B][N][F|N] of this step).  sums4_out (device, optional): the batch sums of the four corrector norms.  S4 steps are one
 * unit (phases must be CORRECTOR | PREDICTOR).  Enqueues on `stream`, no host synchronisation. */
#define GDSS_PHASE_CORRECTOR 1
#define GDSS_PHASE_PREDICTOR 2
int gdss_solver_phase(gdss_ctx* ctx, const gdss_sampler_desc* s, const gdss_step_coef* row, int32_t step, int32_t phases,
                      float* x, float* adj, const float* flags, float* x_mean, float* adj_mean, int32_t B, uint64_t seed,
                      int64_t graph_offset, int64_t global_B, const float* inj_x, const float* inj_adj, float* sums4_out,
                      void* nccl_comm, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- denoising-score-matching loss, forward only -----------------------------------------------
 * Replaces loss_fn of losses.py:37-81 (likelihood_weighting = False) as the trainer's evaluation loop calls it under
 * torch.no_grad() (trainer.py:84-99): flags = node_flags(adj), z = gen_noise from the supplied raw N(0,1) draws
 * (raw_x [B,N,F], raw_adj [B,N,N]), perturbed = mask(mean * y + std * z), both score networks, per-graph
 * reduce_op(square(score * std + z)), batch mean -> loss_out[0] (x), loss_out[1] (adj) on the device.
 * coef = float[B][6]: mean_x, std_x, div_x, mean_adj, std_adj, div_adj from sde.marginal_prob(., t_b); div = std for
 * VP / subVP (score = -net / std), 0 for VE (score = net).  scratch >= gdss_dsm_scratch_floats floats.
 * The optimizer half of trainer.py:66-79 is gdss_adam_ema_step below. */
int64_t gdss_dsm_scratch_floats(const gdss_ctx* ctx, int32_t B);
int gdss_dsm_loss(gdss_ctx* ctx, const float* x, const float* adj, const float* coef, const float* raw_x,
                  const float* raw_adj, int32_t B, int32_t reduce_mean, float* scratch, int64_t scratch_floats,
                  float* loss_out, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- denoising-score-matching training step: both losses AND their parameter gradients -------------------------------
 * Replaces trainer.py:57-65 (loss_fn, loss_x.backward(), loss_adj.backward()) for one batch: the forward of gdss_dsm_loss with
 * every activation kept in `train_workspace`, then the hand-derived reverse sweep of both score networks (DenseGCNConv incl.
 * the gradient through the degree normalisation into the adjacency channels, tanh attention maps, per-edge / node / final
 * MLPs, symmetrisation, masks; layers.py:49-88, :135-153, attention.py:25-51, :93-116, ScoreNetwork_A.py:131-150,
 * ScoreNetwork_X.py:32-46, :75-89).  x / adj / coef / raw_x / raw_adj / reduce_mean as in gdss_dsm_loss.
 * grad_blob_x / grad_blob_a (device, gdss_packed_floats floats each, overwritten): d loss_x / d theta_x and
 * d loss_adj / d theta_adj in the LAYOUT OF THE PACKED BLOBS (padding entries stay 0); parameters a loss never reaches get 0.
 * The context's blobs are read as the current weights (update them in place between steps).  No dead-chain elimination (every
 * parameter gets its gradient); padded nodes and the symmetric half of every per-edge quantity are skipped exactly (their
 * gradient contributions are zero / identical); N <= 128.  loss_out[0..1] on the device.  No host synchronisation.  The node-
 * feature network runs on a side stream of the context next to the adjacency network (fork / join by events on `stream`). */
int64_t gdss_train_workspace_bytes(const gdss_ctx* ctx, int32_t B);
int gdss_dsm_loss_backward(gdss_ctx* ctx, const float* x, const float* adj, const float* coef, const float* raw_x,
                           const float* raw_adj, int32_t B, int32_t reduce_mean, float* grad_blob_x, float* grad_blob_a,
                           float* loss_out, void* train_workspace, int64_t train_workspace_bytes, void* stream);

/* ---- NCCL plumbing (dlopen'ed; the library loads without NCCL) ------------------------- */
int gdss_nccl_unique_id(void* id128);                         /* rank 0: 128-byte id       */
int gdss_nccl_comm_create(const void* id128, int32_t world, int32_t rank, void** comm_out);
int gdss_nccl_comm_destroy(void* comm);
int gdss_nccl_allgather(void* comm, const float* send, float* recv, int64_t count_per_rank, void* stream);
/* in-place sum over the ranks: the data-parallel trainer's gradient all-reduce (replaces nn.DataParallel's reduction,
 * trainer.py:57-79 / utils/loader.py:37-48); pass grad_scale = 1 / world to gdss_adam_ema_step for the batch mean */
int gdss_nccl_allreduce(void* comm, float* buf, int64_t count, void* stream);

/* ---- optimizer half of the training step (trainer.py:66-79) on the flat fp32 arrays of ONE network --------------------
 * clip_grad_norm_(params, grad_clip) (skipped when grad_clip <= 0) -> torch.optim.Adam step (L2 weight decay folded into
 * the gradient, bias correction for the 1-based `step`) -> ExponentialMovingAverage.update (utils/ema.py:18-35; skipped
 * when ema_shadow is NULL; ema_num_updates = updates BEFORE this one, the warm-up decay min(decay, (1+n)/(10+n)) is applied;
 * pass -1 to use ema_decay as it is).  grads are read as grads * grad_scale.  scratch: gdss_optim_scratch_floats() floats.
 * grad_norm_out (optional, device): the total gradient norm before clipping.  active_or_null (device, one byte per
 * parameter): 0 marks parameters the loss never reaches (grad None in torch, e.g. the last AttentionLayer's node MLP of
 * ScoreNetworkA) -- torch.optim.Adam skips them (no weight decay, no moment update), and so does this step; their gradient
 * entries must be 0 (they do not enter the norm).  Two launches, no host synchronisation. */
int64_t gdss_optim_scratch_floats(void);
int gdss_adam_ema_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema_shadow, int64_t n,
                       float lr, float beta1, float beta2, float eps, float weight_decay, float grad_clip, float grad_scale,
                       int32_t step, float ema_decay, int64_t ema_num_updates, const unsigned char* active_or_null,
                       float* scratch, float* grad_norm_out, void* stream);
/* The per-step scalars of that update -- out4 = {lr, 1 - beta1^step, sqrt(1 - beta2^step), 1 - EMA decay of this update} (host,
 * double arithmetic as torch's) -- and the same step reading them from DEVICE memory: the launch no longer depends on the step
 * count, so a whole training step (gdss_dsm_loss_backward + gdss_nccl_allreduce + this) can be captured into one CUDA graph and
 * replayed; the caller copies the four floats to the device before each replay. */
int gdss_adam_step_consts(float lr, float beta1, float beta2, int32_t step, float ema_decay, int64_t ema_num_updates, float* out4);
int gdss_adam_ema_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema_shadow, int64_t n,
                           float beta1, float beta2, float eps, float weight_decay, float grad_clip, float grad_scale,
                           const float* step_consts_dev, const unsigned char* active_or_null, float* scratch,
                           float* grad_norm_out, void* stream);

/* ---- tensor helpers -----------------------------------------------------------------
 * Replace utils/graph_utils.py: mask_x :8-12, mask_adjs :16-29 (3-D or 4-D via C channels),
 * gen_noise :78-86 (raw draw supplied or Philox), pow_tensor :133-142, quantize :90-92,
 * quantize_mol :97-106 (int64 out). */
int gdss_mask_x(float* x, const float* flags, int32_t B, int32_t N, int32_t F, void* stream);
int gdss_mask_adjs(float* adjs, const float* flags, int32_t B, int32_t C, int32_t N, void* stream);
int gdss_gen_noise(float* out, const float* raw_or_null, const float* flags, int32_t B, int32_t N, int32_t F_or_0,
                   int32_t sym, uint64_t seed, uint64_t draw_id, int64_t graph_offset, void* stream);
int gdss_pow_tensor(const float* adj, float* out /* [B,cnum,N,N] */, int32_t B, int32_t N, int32_t cnum, void* stream);
int gdss_quantize(const float* adj, float* out, float thr, int64_t numel, void* stream);
int gdss_quantize_mol(const float* adj, int64_t* out, int64_t numel, void* stream);
/* Post-sampling decode of a molecule batch in one pass (replaces sampler.py:131-138: quantize_mol, the bond-class
 * remap 0,1,2,3 -> 3,0,1,2, x > 0.5 and the virtual-atom channel) with compact int8 outputs:
 * atom [B][N][F+1], bond [B][N][N] (class index; one_hot(bond, 4).permute(0,3,1,2) is the reference's adj). */
int gdss_decode_mol(const float* x, const float* adj, int8_t* atom, int8_t* bond, int32_t B, int32_t N, int32_t F,
                    void* stream);
/* init_flags (utils/graph_utils.py:66-74) on the device: node_count_cdf = float[N+1], inclusive cumulative counts of
 * the training graphs' node numbers 0..N; graph b gets its first n_b flags set, n_b drawn from that histogram with the
 * library's Philox stream keyed by (seed, graph_offset + b): independent of how the batch is sharded. */
int gdss_init_flags(const float* node_count_cdf, float* flags, int32_t B, int32_t N, uint64_t seed, int64_t graph_offset,
                    void* stream);

/* ---- introspection for tests / profiling ---------------------------------------------- */
/* number of kernel launches the library has enqueued since process start (all contexts) */
int64_t gdss_launch_count(void);
/* per-kernel-class device time of everything enqueued (eagerly, i.e. use_graph == 0) between
 * gdss_prof_begin and gdss_prof_end: CUDA events recorded after every launch on its own stream.
 * ms / counts have n >= 16 entries indexed by kernel class; gdss_prof_name names a class. */
int gdss_prof_begin(void* stream);
int gdss_prof_end(float* ms, int64_t* counts, int32_t n);
const char* gdss_prof_name(int32_t kernel_class);
/* workspace layout of the last plan: byte offsets of named regions (for parity tests that look
 * at intermediates).  Returns <0 if name unknown. */
int64_t gdss_workspace_offset(const gdss_ctx* ctx, int32_t B, const char* region);
/* which kernel path a context selected: key = "fused" (1: per-graph fused kernel, 0: tiled general path),
 * "fused_fusable" (the networks' shapes are inside the fused kernel's limits), "fused_smem_bytes" (shared memory
 * the per-graph kernel needs for N nodes; -1 if its regions cannot be planned), "bucket_smem_bytes:<NB>:<0|1>" (the same for a
 * size bucket of at most NB nodes, with / without the layer weights staged), "smem_optin", "use_tc", "use_mma".
 * Returns <0 if the key is unknown. */
int64_t gdss_ctx_query(const gdss_ctx* ctx, const char* key);

#ifdef __cplusplus
}
#endif
#endif /* GDSS_B200_H */
