// Weight packing: reference state_dict (utils/loader.py:186-196 load_state_dict) -> one fp32 blob
// per network with a fixed layout (gdss_internal.h NetPlan).  Host code only.
#include <math.h>
#include <string.h>

#include "gdss_internal.h"

namespace gdss {

static inline int64_t take(int64_t& cur, int64_t n) {
  int64_t o = cur;
  cur += (n + 3) & ~int64_t(3);          // keep every tensor 16-byte aligned for float4 loads
  return o;
}

// Shape rules: models/ScoreNetwork_A.py:113-127, models/ScoreNetwork_X.py:17-30, :57-72,
// models/attention.py:77-91 (verified against all seven shipped checkpoints).
int plan_net(const gdss_net_desc* d, int N, NetPlan* p) {
  if (!d || !p) return GDSS_E_BADARG;
  memset(p, 0, sizeof(*p));
  p->type = d->model_type;
  p->F = d->max_feat_num;
  p->N = N;
  p->nhid = d->nhid;
  p->depth = d->depth;
  p->nlin = d->num_linears;
  p->c_init = d->c_init;
  p->c_hid = d->c_hid;
  p->c_final = d->c_final;
  p->adim = d->adim;
  p->heads = d->num_heads > 0 ? d->num_heads : 4;
  if (p->F <= 0 || N <= 0 || p->nhid <= 0 || p->depth <= 0 || p->depth > GDSS_MAX_LAYERS) return GDSS_E_BADARG;
  int64_t cur = 0;
  int nt = 0;
  if (p->type == GDSS_NET_X) {
    for (int l = 0; l < p->depth; ++l) {
      int fin = l == 0 ? p->F : p->nhid;
      p->gw[l] = take(cur, (int64_t)((fin + 3) & ~3) * p->nhid);
      p->gb[l] = take(cur, p->nhid);
      nt += 2;
    }
    p->fdim = p->F + p->depth * p->nhid;
    p->fout = p->F;
  } else if (p->type == GDSS_NET_X_GMH || p->type == GDSS_NET_A) {
    if (p->nlin < 2 || p->nlin > GDSS_MAX_LIN) return GDSS_E_UNSUPPORTED;
    if (p->c_init < 1 || p->c_hid < 1 || p->c_final < 1) return GDSS_E_BADARG;
    if (p->c_init > GDSS_MAX_CH || p->c_hid > GDSS_MAX_CH || p->c_final > GDSS_MAX_CH) return GDSS_E_UNSUPPORTED;
    if (p->adim <= 0 || p->nhid % p->heads || p->adim % p->heads) return GDSS_E_BADARG;
    int chans = p->c_init;
    for (int l = 0; l < p->depth; ++l) {
      LayerPlan& L = p->layers[l];
      L.c_in = l == 0 ? p->c_init : p->c_hid;
      L.c_out = l == 0 ? p->c_hid : (l == p->depth - 1 ? p->c_final : p->c_hid);
      L.f_in = l == 0 ? p->F : p->nhid;
      L.attn = l == 0 ? p->nhid : p->adim;
      L.nhid = p->nhid;
      L.hid = 2 * (L.c_in > L.c_out ? L.c_in : L.c_out);
      L.nlin = p->nlin;
      L.qkvw = 2 * L.attn + L.nhid;
      L.f_in_p = (L.f_in + 3) & ~3;
      L.wqkv = take(cur, (int64_t)L.c_in * L.f_in_p * L.qkvw);
      L.bqkv = take(cur, (int64_t)L.c_in * L.qkvw);
      nt += 6 * L.c_in;
      for (int i = 0; i < L.nlin; ++i) {
        int in = i == 0 ? 2 * L.c_in : L.hid, out = i == L.nlin - 1 ? L.c_out : L.hid;
        L.ew[i] = take(cur, (int64_t)in * out);
        L.eb[i] = take(cur, out);
        nt += 2;
      }
      L.nw[0] = take(cur, (int64_t)L.c_in * L.nhid * L.hid);
      L.nb[0] = take(cur, L.hid);
      L.nw[1] = take(cur, (int64_t)L.hid * L.nhid);
      L.nb[1] = take(cur, L.nhid);
      nt += 4;
      chans += L.c_out;
      L.qk_live = L.v_live = (1u << L.c_in) - 1u;
      L.node_live = 1;
      L.edge_const = 0;
    }
    if (p->type == GDSS_NET_A) {
      p->fdim = p->c_hid * (p->depth - 1) + p->c_final + p->c_init;
      if (p->fdim != chans) return GDSS_E_UNSUPPORTED;   // the reference itself fails to run in this case
      p->fout = 1;
    } else {
      p->fdim = p->F + p->depth * p->nhid;
      p->fout = p->F;
    }
  } else {
    return GDSS_E_UNSUPPORTED;
  }
  p->fdp = (p->fdim + 3) & ~3;
  p->Hp = (2 * p->fdim + 3) & ~3;
  p->foutp = (p->fout + 3) & ~3;
  int dims[4] = {p->fdp, p->Hp, p->Hp, p->foutp};
  for (int i = 0; i < 3; ++i) {
    p->fw[i] = take(cur, (int64_t)dims[i] * dims[i + 1]);
    p->fb[i] = take(cur, dims[i + 1]);
    nt += 2;
  }
  p->n_tensors = nt;
  p->total = cur;
  return GDSS_OK;
}

// nn.Linear weight [out][in] -> [in][out] with row stride ld (>= out; the padding stays zero)
static void put_linear_t(float* dst, const float* src, int in, int out, int ld = 0) {
  if (ld <= 0) ld = out;
  for (int o = 0; o < out; ++o)
    for (int i = 0; i < in; ++i) dst[(int64_t)i * ld + o] = src[(int64_t)o * in + i];
}

// ---------------------------------------------------------------------------------------
// Dead sub-network elimination.  The shipped checkpoints were trained with weight decay; whole weight blocks of them
// have decayed into the fp32 denormal range (|w| ~ 1e-39: e.g. every gnn_q / gnn_k of layer 0 and the complete last
// AttentionLayer of ScoreNetworkA in gdss_zinc250k_v2.pth; the EMA shadows ENZYMES / grid sample with lag behind at
// 1e-14 .. 1e-30).  A block below GDSS_DEAD_EPS = 1e-12 moves sums of O(1e-2 .. 1) terms by < 1e-9 relative -- 1/60 of an
// fp32 ulp -- so the chain that only feeds it is not evaluated:
//   * Q or K convolution dead (weights and bias)  -> the attention map is tanh(0) = 0 exactly (attention.py:41-49)
//   * row c of mlp.linears.0 dead                 -> map c is never read (attention.py:109-110)
//   * rows of multi_channel.linears.0 for V_c dead, or x_out unread downstream -> V_c is never read (attention.py:106)
//   * every row of mlp.linears.0 dead             -> the layer's channels are the constant mlp(0) (times the flag mask)
// Liveness of x_out is propagated from the last layer backwards (ScoreNetworkA reads x only through the next layer's
// convolutions, ScoreNetwork_A.py:136-139; ScoreNetworkX_GMH concatenates every x, ScoreNetwork_X.py:79-84).
// Random-init weights have no dead blocks and are evaluated in full.
// ---------------------------------------------------------------------------------------
#define GDSS_DEAD_EPS 1e-12f

static bool block_dead(const float* w, int64_t rows, int64_t cols, int64_t ld) {
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t c = 0; c < cols; ++c)
      if (!(fabsf(w[r * ld + c]) < GDSS_DEAD_EPS)) return false;      // NaN counts as live
  return true;
}

static float elu_host(float v) { return v > 0.f ? v : expm1f(v); }

int analyze_liveness(NetPlan* p, const float* blob) {
  if (!p || !blob || p->type == GDSS_NET_X) return 0;
  const bool isA = p->type == GDSS_NET_A;
  int off = 0;
  bool x_needed = !isA;                      // is the x_out of the layer being visited read by a later layer / the final MLP?
  for (int l = p->depth - 1; l >= 0; --l) {
    LayerPlan& L = p->layers[l];
    const bool last = l == p->depth - 1;
    const bool need_edge = isA || !last, need_node = isA ? (!last && x_needed) : true;
    const int C = L.c_in, h = L.nhid, A = L.attn;
    const float* W1 = blob + L.ew[0];                       // [2C][hid]
    L.edge_const = (need_edge && block_dead(W1, 2 * C, L.hid, L.hid)) ? 1 : 0;
    if (L.edge_const) {
      float hbuf[2][64], *cur = hbuf[0], *nxt = hbuf[1];
      for (int o = 0; o < L.hid; ++o) cur[o] = elu_host(blob[L.eb[0] + o]);
      for (int i = 1; i < L.nlin; ++i) {
        const int out = i == L.nlin - 1 ? L.c_out : L.hid;
        for (int o = 0; o < out; ++o) {
          float s = blob[L.eb[i] + o];
          for (int k = 0; k < L.hid; ++k) s = fmaf(cur[k], blob[L.ew[i] + (int64_t)k * out + o], s);
          nxt[o] = i == L.nlin - 1 ? s : elu_host(s);
        }
        float* t = cur; cur = nxt; nxt = t;
      }
      for (int o = 0; o < L.c_out && o < GDSS_MAX_CH; ++o) L.edge_cst[o] = cur[o];
      ++off;
    }
    uint32_t qk = 0, v = 0;
    for (int c = 0; c < C; ++c) {
      const float* Wc = blob + L.wqkv + (int64_t)c * L.f_in_p * L.qkvw;
      const float* bc = blob + L.bqkv + (int64_t)c * L.qkvw;
      const bool qdead = block_dead(Wc, L.f_in, A, L.qkvw) && block_dead(bc, 1, A, L.qkvw);
      const bool kdead = block_dead(Wc + A, L.f_in, A, L.qkvw) && block_dead(bc + A, 1, A, L.qkvw);
      const bool vdead = block_dead(Wc + 2 * A, L.f_in, h, L.qkvw) && block_dead(bc + 2 * A, 1, h, L.qkvw);
      const bool mapdead = block_dead(W1 + (int64_t)c * L.hid, 1, L.hid, L.hid);
      const bool vblkdead = block_dead(blob + L.nw[0] + (int64_t)c * h * L.hid, h, L.hid, L.hid);
      if (need_edge && !L.edge_const && !qdead && !kdead && !mapdead) qk |= 1u << c;
      if (need_node && !vdead && !vblkdead) v |= 1u << c;
    }
    const uint32_t all = (1u << C) - 1u;
    const uint32_t qk_base = need_edge ? all : 0u, v_base = need_node ? all : 0u;
    for (uint32_t m = (qk_base & ~qk) | (v_base & ~v); m; m &= m - 1) ++off;
    L.qk_live = qk;
    L.v_live = v;
    L.node_live = need_node ? 1 : 0;
    if (isA && !last && !need_node) ++off;
    // x_{l-1} is read by this layer's convolutions only
    x_needed = isA ? (qk != 0 || v != 0) : true;
  }
  return off;
}

struct Cursor {
  const float* const* t;
  const int64_t* numel;
  int n, pos;
  bool bad;
  const float* next(int64_t expect) {
    if (pos >= n || numel[pos] != expect || !t[pos]) {
      bad = true;
      return nullptr;
    }
    return t[pos++];
  }
};

}  // namespace gdss

using namespace gdss;

extern "C" int64_t gdss_packed_floats(const gdss_net_desc* desc, int32_t n_nodes) {
  NetPlan p;
  int rc = plan_net(desc, n_nodes, &p);
  return rc < 0 ? rc : p.total;
}

extern "C" int32_t gdss_num_state_tensors(const gdss_net_desc* desc) {
  NetPlan p;
  int rc = plan_net(desc, desc && desc->max_node_num > 0 ? desc->max_node_num : 1, &p);
  return rc < 0 ? rc : p.n_tensors;
}

extern "C" int gdss_analyze_weights(const gdss_net_desc* desc, int32_t n_nodes, const float* host_blob, int32_t* out,
                                    int32_t max_layers) {
  NetPlan p;
  int rc = plan_net(desc, n_nodes, &p);
  if (rc < 0) return rc;
  if (!host_blob || !out || max_layers < p.depth) return GDSS_E_BADARG;
  const int off = analyze_liveness(&p, host_blob);
  for (int l = 0; l < p.depth; ++l) {
    const LayerPlan& L = p.layers[l];
    const bool att = p.type != GDSS_NET_X;
    out[4 * l + 0] = att ? (int32_t)L.qk_live : 0;
    out[4 * l + 1] = att ? (int32_t)L.v_live : 0;
    out[4 * l + 2] = att ? L.node_live : 1;
    out[4 * l + 3] = att ? L.edge_const : 0;
  }
  return off;
}

extern "C" int gdss_pack_weights(const gdss_net_desc* desc, int32_t n_nodes, const float* const* tensors,
                                 const int64_t* numel, int32_t n_tensors, float* blob, int64_t blob_floats) {
  NetPlan p;
  int rc = plan_net(desc, n_nodes, &p);
  if (rc < 0) return rc;
  if (!tensors || !numel || !blob || n_tensors != p.n_tensors || blob_floats < p.total) return GDSS_E_BADARG;
  memset(blob, 0, sizeof(float) * p.total);
  Cursor c{tensors, numel, n_tensors, 0, false};
  if (p.type == GDSS_NET_X) {
    for (int l = 0; l < p.depth; ++l) {
      int fin = l == 0 ? p.F : p.nhid;
      const float* w = c.next((int64_t)fin * p.nhid);     // DenseGCNConv.weight is already [in][out]
      const float* b = c.next(p.nhid);
      if (c.bad) return GDSS_E_BADARG;
      memcpy(blob + p.gw[l], w, sizeof(float) * fin * p.nhid);
      memcpy(blob + p.gb[l], b, sizeof(float) * p.nhid);
    }
  } else {
    for (int l = 0; l < p.depth; ++l) {
      const LayerPlan& L = p.layers[l];
      for (int ch = 0; ch < L.c_in; ++ch) {
        // order in the state_dict: gnn_q.{weight,bias}, gnn_k.{weight,bias}, gnn_v.{weight,bias}
        int widths[3] = {L.attn, L.attn, L.nhid};
        int col = 0;
        for (int m = 0; m < 3; ++m) {
          const float* w = c.next((int64_t)L.f_in * widths[m]);
          const float* b = c.next(widths[m]);
          if (c.bad) return GDSS_E_BADARG;
          float* W = blob + L.wqkv + (int64_t)ch * L.f_in_p * L.qkvw;
          for (int f = 0; f < L.f_in; ++f)
            for (int o = 0; o < widths[m]; ++o) W[(int64_t)f * L.qkvw + col + o] = w[(int64_t)f * widths[m] + o];
          float* Bv = blob + L.bqkv + (int64_t)ch * L.qkvw;
          for (int o = 0; o < widths[m]; ++o) Bv[col + o] = b[o];
          col += widths[m];
        }
      }
      for (int i = 0; i < L.nlin; ++i) {
        int in = i == 0 ? 2 * L.c_in : L.hid, out = i == L.nlin - 1 ? L.c_out : L.hid;
        const float* w = c.next((int64_t)in * out);
        const float* b = c.next(out);
        if (c.bad) return GDSS_E_BADARG;
        put_linear_t(blob + L.ew[i], w, in, out);
        memcpy(blob + L.eb[i], b, sizeof(float) * out);
      }
      for (int i = 0; i < 2; ++i) {
        int in = i == 0 ? L.c_in * L.nhid : L.hid, out = i == 0 ? L.hid : L.nhid;
        const float* w = c.next((int64_t)in * out);
        const float* b = c.next(out);
        if (c.bad) return GDSS_E_BADARG;
        put_linear_t(blob + L.nw[i], w, in, out);
        memcpy(blob + L.nb[i], b, sizeof(float) * out);
      }
    }
  }
  int dims[4] = {p.fdim, 2 * p.fdim, 2 * p.fdim, p.fout};
  int lds[3] = {p.Hp, p.Hp, p.foutp};
  for (int i = 0; i < 3; ++i) {
    const float* w = c.next((int64_t)dims[i] * dims[i + 1]);
    const float* b = c.next(dims[i + 1]);
    if (c.bad) return GDSS_E_BADARG;
    put_linear_t(blob + p.fw[i], w, dims[i], dims[i + 1], lds[i]);
    memcpy(blob + p.fb[i], b, sizeof(float) * dims[i + 1]);
  }
  if (c.pos != n_tensors) return GDSS_E_BADARG;
  return GDSS_OK;
}
