// Weight packing: reference state_dict (utils/loader.py:186-196 load_state_dict) -> one fp32 blob
// per network with a fixed layout (gdss_internal.h NetPlan).  Host code only.
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
