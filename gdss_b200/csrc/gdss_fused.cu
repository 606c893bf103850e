// Fused small-graph path (N <= 64): one CTA evaluates BOTH score networks' layer stacks for one graph with
// every intermediate (normalised adjacency channels, GCN products, Q|K|V, attention maps, hidden units) in
// shared memory; only the A-network's per-pixel channel stack leaves the SM, in a pixel-major layout that the
// batched final-MLP kernel (ScoreNetwork_A.py:141-148) consumes in 128-pixel tiles over the whole batch.
//
// Data layout
//   pixel p of graph b = strict upper-triangle entry (i < j < n_b), row-major; n_b = neff[b].  All per-edge
//   quantities are symmetric in (i, j) (attention.py:49, :113) and the diagonal never reaches an output
//   (DenseGCNConv overwrites it with 1, layers.py:72-74; the final score zeroes it, ScoreNetwork_A.py:145).
//   global pixel index gp = poff[b] + p;   stackT[c][gp]  (c = channel of the concatenated stack)
//   gmap[gp] = b << 12 | i << 6 | j        (-1 on the padding of the last 128-pixel tile)
#include <stdio.h>
#include <string.h>

#include "gdss_common.cuh"
#include "gdss_internal.h"

namespace gdss {

#define LAUNCH_CHECK()                          \
  do {                                          \
    cudaError_t e__ = cudaGetLastError();       \
    if (e__ != cudaSuccess) return (int)e__;    \
  } while (0)

DEVINL float tanh_fast(float v) {      // no clamp needed: ex2 -> inf gives 1, -> 0 gives -1
  float e = __expf(2.f * v);
  return 1.f - __fdividef(2.f, e + 1.f);
}

// ---------------------------------------------------------------------------------------
// pixel bookkeeping (depends on the flags only: once per sampler call)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) tri_scan_kernel(const int* __restrict__ neff, int B, int* __restrict__ poff,
                                                        int* __restrict__ meta, int* __restrict__ gmap) {
  __shared__ int part[1024];
  const int t = threadIdx.x;
  const int chunk = (B + 1023) / 1024;
  const int lo = min(B, t * chunk), hi = min(B, lo + chunk);
  int s = 0;
  for (int b = lo; b < hi; ++b) {
    int n = neff[b];
    s += n * (n - 1) / 2;
  }
  part[t] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    int v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int base = part[t] - s;
  for (int b = lo; b < hi; ++b) {
    poff[b] = base;
    int n = neff[b];
    base += n * (n - 1) / 2;
  }
  if (t == 1023) {
    const int tot = part[1023];
    poff[B] = tot;
    meta[0] = tot;
    meta[1] = (tot + 127) / 128;
    for (int k = tot; k < ((tot + 127) / 128) * 128; ++k) gmap[k] = -1;
  }
}

__global__ void __launch_bounds__(64) tri_map_kernel(const int* __restrict__ neff, const int* __restrict__ poff,
                                                     int* __restrict__ gmap) {
  const int b = blockIdx.x, n = neff[b], base = poff[b];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int rs = i * (2 * n - i - 1) / 2;
    for (int j = i + 1; j < n; ++j) gmap[base + rs + j - i - 1] = (b << 12) | (i << 6) | j;
  }
}

// ---------------------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------------------
#define FUSED_MAX_LAYERS 8
#define RMAX 8                 // rows per register-blocked item

struct FLayer {
  int c_in, c_out, f_in, attn, nlin, qkvw, need_edge, need_node;
  const float *wqkv, *bqkv, *ew[3], *eb[3], *nw0, *nb0, *nw1, *nb1;
};

struct FNet {
  int type, depth, F, nhid, heads, fdim, fout, c_init;
  FLayer L[FUSED_MAX_LAYERS];
  const float* gw[FUSED_MAX_LAYERS];
  const float* gb[FUSED_MAX_LAYERS];
  const float *fw[3], *fb[3];
};

// float offsets into dynamic shared memory
struct FSmem {
  int fl, dinv, pix, x0, xb, xc, xs, h1, wts, p0, pa, r1, qk, r2, total;
  int LD, XS, QKS, PM, NBP, XSS;
};

struct FusedArgs {
  FNet a, x;
  int has_a, has_x;
  const float *x_in, *adj, *flags;
  const int* neff;
  float* raw_x;
  float* stackT;
  const int* poff;
  int64_t pstride;
  int N, F;
  FSmem sl;
};

// ---------------------------------------------------------------------------------------
// mm_rows: out[c][i][o] = act( (bias[c][o] + sum_k in[c][i][k] * W[c][k][o]) * rowscale[i] )
// Register-blocked: one item = 4 consecutive output columns x up to RMAX rows; the 4x4 weight block of a
// k-chunk is loaded once (global, L2-resident) and reused over the rows, the input rows come from shared
// memory as float4.  `in` rows must be readable (and finite) up to K rounded up to 4.
// Output columns < split go to out0, the others to out1 (column index minus split).
// ---------------------------------------------------------------------------------------
enum { ACT_NONE = 0, ACT_ELU = 1, ACT_TANH = 2, ACT_TANH2 = 3 };

template <int ACT>
DEVINL void mm_rows(const float* in, int in_cs, int ldin, int K, const float* __restrict__ W, int w_cs, int ldw,
                    const float* __restrict__ bias, int b_cs, int C, int n, int o_lo, int o_hi, int OUTN, float* out0,
                    int o0_cs, int ld0, int split, float* out1, int o1_cs, int ld1, const float* rowscale) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ogn = (o_hi - o_lo + 3) >> 2;
  const int pairs = C * ogn;
  if (pairs <= 0 || n <= 0) return;
  // rows are dealt round-robin to IP row-partitions; pick IP to minimise (rounds x rows per item)
  int IP = (n + RMAX - 1) / RMAX, best = 1 << 30;
  for (int rows = RMAX; rows >= 1; --rows) {
    const int ip = (n + rows - 1) / rows;
    const int rounds = (pairs * ip + nt - 1) / nt;
    const int cost = rounds * (rows * 5 + 4);      // ~ weight loads + rows * (lds + 16 fma)
    if (cost < best) { best = cost; IP = ip; }
  }
  const int K4 = (K + 3) & ~3;
  const bool vecw = ((ldw & 3) == 0);
  for (int item = tid; item < pairs * IP; item += nt) {
    const int ip = item / pairs, pr = item - ip * pairs;
    const int c = pr / ogn, og = pr - c * ogn;
    const int o = o_lo + 4 * og;
    const float* Wc = W + (int64_t)c * w_cs;
    const float* inc = in + c * in_cs;
    const int rcnt = (n - ip + IP - 1) / IP;       // rows of this item: ip, ip + IP, ...
    float acc[RMAX][4];
    {
      float bv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) bv[q] = (o + q < OUTN) ? __ldg(bias + c * b_cs + o + q) : 0.f;
#pragma unroll
      for (int r = 0; r < RMAX; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = bv[q];
    }
    for (int k0 = 0; k0 < K4; k0 += 4) {
      float w[4][4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = k0 + kk;
        if (k < K) {
          if (vecw && o + 3 < OUTN) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(Wc + (int64_t)k * ldw + o));
            w[kk][0] = t.x; w[kk][1] = t.y; w[kk][2] = t.z; w[kk][3] = t.w;
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) w[kk][q] = (o + q < OUTN) ? __ldg(Wc + (int64_t)k * ldw + o + q) : 0.f;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) w[kk][q] = 0.f;
        }
      }
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        if (r >= rcnt) break;
        const int i = ip + r * IP;
        const float4 a = *reinterpret_cast<const float4*>(inc + i * ldin + k0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[r][q] = fmaf(a.x, w[0][q], acc[r][q]);
          acc[r][q] = fmaf(a.y, w[1][q], acc[r][q]);
          acc[r][q] = fmaf(a.z, w[2][q], acc[r][q]);
          acc[r][q] = fmaf(a.w, w[3][q], acc[r][q]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      const int i = ip + r * IP;
      if (i >= n) continue;
      const float rs = rowscale ? rowscale[i] : 1.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int oc = o + q;
        if (oc >= o_hi) continue;                 // columns in [OUTN, o_hi) are padding: exactly act(0) = 0
        float v = acc[r][q] * rs;
        if (ACT == ACT_ELU) v = elu_f(v);
        if (ACT == ACT_TANH) v = tanh_fast(v);
        if (ACT == ACT_TANH2) v = tanh_fast(tanh_fast(v));
        if (oc < split) out0[c * o0_cs + i * ld0 + oc] = v;
        else out1[c * o1_cs + i * ld1 + (oc - split)] = v;
      }
    }
  }
}

DEVINL int tri_index(int i, int j, int n) {      // i != j
  const int lo = min(i, j), hi = max(i, j);
  return lo * (2 * n - lo - 1) / 2 + hi - lo - 1;
}

// dense normalised adjacency of C tri planes: D^-1/2 (P with unit diagonal) D^-1/2   (layers.py:72-79)
DEVINL void build_norm_adj(const float* pin, int PM, int C, int n, int LD, int NBP, float* dense, float* dinv) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nn = n * n;
  for (int idx = tid; idx < C * nn; idx += nt) {
    const int c = idx / nn, r = idx - c * nn;
    const int i = r / n, j = r - i * n;
    dense[c * n * LD + i * LD + j] = (i == j) ? 1.f : pin[c * PM + tri_index(i, j, n)];
  }
  __syncthreads();
  for (int idx = tid; idx < C * n; idx += nt) {
    const int c = idx / n, i = idx - c * n;
    const float* row = dense + c * n * LD + i * LD;
    float s = 0.f;
    for (int j = 0; j < n; ++j) s += row[j];
    dinv[c * NBP + i] = 1.0f / sqrtf(fmaxf(s, 1.0f));
  }
  __syncthreads();
  for (int idx = tid; idx < C * nn; idx += nt) {
    const int c = idx / nn, r = idx - c * nn;
    const int i = r / n, j = r - i * n;
    float* e = dense + c * n * LD + i * LD + j;
    *e = (dinv[c * NBP + i] * *e) * dinv[c * NBP + j];
  }
  __syncthreads();
}

// AX[c][i][0..fp) = sum_j dense[c][i][j] * x[j][0..fp)     (x rows zero-padded to fp = f rounded up to 4)
DEVINL void adj_times_x(const float* dense, int LD, const float* x, int ldx, int C, int n, int fp, float* ax) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int fg = fp >> 2;
  for (int item = tid; item < C * n * fg; item += nt) {
    const int g = item % fg, ci = item / fg;
    const int c = ci / n, i = ci - c * n;
    const float* row = dense + c * n * LD + i * LD;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < n; ++j) {
      const float a = row[j];
      const float4 xv = *reinterpret_cast<const float4*>(x + j * ldx + 4 * g);
      acc.x = fmaf(a, xv.x, acc.x);
      acc.y = fmaf(a, xv.y, acc.y);
      acc.z = fmaf(a, xv.z, acc.z);
      acc.w = fmaf(a, xv.w, acc.w);
    }
    *reinterpret_cast<float4*>(ax + (c * n + i) * fp + 4 * g) = acc;
  }
}

// stage the per-edge MLP of one AttentionLayer (attention.py:89, layers.py:96-153) into shared memory:
//   W1 [2C][16] b1[16] | (W2 [16][16] b2[16]) | W3 [16][8] b3[8]      (hidden width 16, outputs padded to 8)
DEVINL void stage_edge_weights(const FLayer& L, float* w) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int in1 = 2 * L.c_in;
  int cur = 0;
  for (int idx = tid; idx < in1 * 16; idx += nt) w[cur + idx] = __ldg(L.ew[0] + idx);
  cur += in1 * 16;
  for (int idx = tid; idx < 16; idx += nt) w[cur + idx] = __ldg(L.eb[0] + idx);
  cur += 16;
  if (L.nlin == 3) {
    for (int idx = tid; idx < 256; idx += nt) w[cur + idx] = __ldg(L.ew[1] + idx);
    cur += 256;
    for (int idx = tid; idx < 16; idx += nt) w[cur + idx] = __ldg(L.eb[1] + idx);
    cur += 16;
  }
  const float* wl = L.ew[L.nlin - 1];
  const float* bl = L.eb[L.nlin - 1];
  for (int idx = tid; idx < 128; idx += nt) {
    const int u = idx >> 3, o = idx & 7;
    w[cur + idx] = o < L.c_out ? __ldg(wl + u * L.c_out + o) : 0.f;
  }
  cur += 128;
  for (int idx = tid; idx < 8; idx += nt) w[cur + idx] = idx < L.c_out ? __ldg(bl + idx) : 0.f;
}

// attention maps (attention.py:35-49): M[c][p] = mean over heads and over the two orientations of
// tanh(Q_h . K_h / sqrt(out_dim))
DEVINL void attention_maps(const float* qk, int QKS, const unsigned* pix, int C, int n, int P, int PM, int attn,
                           int heads, float scale, float* M) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int hd = attn / heads;
  const float ms = 0.5f / (float)heads;
  for (int item = tid; item < C * P; item += nt) {
    const int c = item / P, p = item - c * P;
    const unsigned ij = pix[p];
    const int i = ij >> 8, j = ij & 255;
    const float* qi = qk + (c * n + i) * QKS;
    const float* qj = qk + (c * n + j) * QKS;
    const float* ki = qi + attn;
    const float* kj = qj + attn;
    float m = 0.f;
    for (int h = 0; h < heads; ++h) {
      float s1 = 0.f, s2 = 0.f;
      for (int d = 0; d < hd; d += 4) {
        const float4 a1 = *reinterpret_cast<const float4*>(qi + h * hd + d);
        const float4 b1 = *reinterpret_cast<const float4*>(kj + h * hd + d);
        const float4 a2 = *reinterpret_cast<const float4*>(qj + h * hd + d);
        const float4 b2 = *reinterpret_cast<const float4*>(ki + h * hd + d);
        s1 = fmaf(a1.x, b1.x, s1); s1 = fmaf(a1.y, b1.y, s1); s1 = fmaf(a1.z, b1.z, s1); s1 = fmaf(a1.w, b1.w, s1);
        s2 = fmaf(a2.x, b2.x, s2); s2 = fmaf(a2.y, b2.y, s2); s2 = fmaf(a2.z, b2.z, s2); s2 = fmaf(a2.w, b2.w, s2);
      }
      m += tanh_fast(s1 * scale) + tanh_fast(s2 * scale);
    }
    M[c * PM + p] = m * ms;
  }
}

// per-pixel MLP of one AttentionLayer (attention.py:109-114): two lanes per pixel, 8 hidden units each.
//   in = [M_0..M_{C-1}, P_0..P_{C-1}] -> 16 -> (16) -> c_out ; out = 2 * mlp * flag_i * flag_j
// pout may alias pin (each pixel is read completely before it is written).
template <int NLIN>
DEVINL void edge_mlp(const float* M, const float* pin, int PM, const unsigned* pix, const float* fl, const float* w,
                     int C, int P, int c_out, float* pout, float* gout, int64_t gstride) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int sub = tid & 1;
  const int in1 = 2 * C;
  const float* W1 = w;
  const float* b1 = W1 + in1 * 16;
  const float* W2 = b1 + 16;
  const float* b2 = W2 + 256;
  const float* W3 = NLIN == 3 ? b2 + 16 : b1 + 16;
  const float* b3 = W3 + 128;
  const int total = 2 * P;
  for (int base = tid - (tid & 31); base < total; base += nt) {       // warp-uniform trip count
    const int idx = base + (tid & 31);
    const bool valid = idx < total;
    const int p = valid ? (idx >> 1) : (P - 1);
    float h[8];
    {
      const float4 ba = *reinterpret_cast<const float4*>(b1 + 8 * sub);
      const float4 bb = *reinterpret_cast<const float4*>(b1 + 8 * sub + 4);
      h[0] = ba.x; h[1] = ba.y; h[2] = ba.z; h[3] = ba.w; h[4] = bb.x; h[5] = bb.y; h[6] = bb.z; h[7] = bb.w;
    }
    for (int k = 0; k < in1; ++k) {
      const float v = k < C ? M[k * PM + p] : pin[(k - C) * PM + p];
      const float4 wa = *reinterpret_cast<const float4*>(W1 + k * 16 + 8 * sub);
      const float4 wb = *reinterpret_cast<const float4*>(W1 + k * 16 + 8 * sub + 4);
      h[0] = fmaf(v, wa.x, h[0]); h[1] = fmaf(v, wa.y, h[1]); h[2] = fmaf(v, wa.z, h[2]); h[3] = fmaf(v, wa.w, h[3]);
      h[4] = fmaf(v, wb.x, h[4]); h[5] = fmaf(v, wb.y, h[5]); h[6] = fmaf(v, wb.z, h[6]); h[7] = fmaf(v, wb.w, h[7]);
    }
    __syncwarp();                     // both lanes of the pixel have read their inputs (in-place update below)
    float hf[16];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float own = elu_f(h[u]);
      const float oth = __shfl_xor_sync(0xffffffffu, own, 1);
      hf[u] = sub ? oth : own;
      hf[8 + u] = sub ? own : oth;
    }
    if (NLIN == 3) {
      const float4 ba = *reinterpret_cast<const float4*>(b2 + 8 * sub);
      const float4 bb = *reinterpret_cast<const float4*>(b2 + 8 * sub + 4);
      h[0] = ba.x; h[1] = ba.y; h[2] = ba.z; h[3] = ba.w; h[4] = bb.x; h[5] = bb.y; h[6] = bb.z; h[7] = bb.w;
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float v = hf[u];
        const float4 wa = *reinterpret_cast<const float4*>(W2 + u * 16 + 8 * sub);
        const float4 wb = *reinterpret_cast<const float4*>(W2 + u * 16 + 8 * sub + 4);
        h[0] = fmaf(v, wa.x, h[0]); h[1] = fmaf(v, wa.y, h[1]); h[2] = fmaf(v, wa.z, h[2]); h[3] = fmaf(v, wa.w, h[3]);
        h[4] = fmaf(v, wb.x, h[4]); h[5] = fmaf(v, wb.y, h[5]); h[6] = fmaf(v, wb.z, h[6]); h[7] = fmaf(v, wb.w, h[7]);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float own = elu_f(h[u]);
        const float oth = __shfl_xor_sync(0xffffffffu, own, 1);
        hf[u] = sub ? oth : own;
        hf[8 + u] = sub ? own : oth;
      }
    }
    float4 o4 = *reinterpret_cast<const float4*>(b3 + 4 * sub);
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const float v = hf[u];
      const float4 wv = *reinterpret_cast<const float4*>(W3 + u * 8 + 4 * sub);
      o4.x = fmaf(v, wv.x, o4.x); o4.y = fmaf(v, wv.y, o4.y); o4.z = fmaf(v, wv.z, o4.z); o4.w = fmaf(v, wv.w, o4.w);
    }
    if (valid) {
      const unsigned ij = pix[p];
      const float fm = 2.f * fl[ij >> 8] * fl[ij & 255];        // (S + S^T) then mask_adjs
      const float ov[4] = {o4.x * fm, o4.y * fm, o4.z * fm, o4.w * fm};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int o = 4 * sub + q;
        if (o < c_out) {
          if (pout) pout[o * PM + p] = ov[q];
          if (gout) gout[(int64_t)o * gstride + p] = ov[q];
        }
      }
    }
  }
}

// final node MLP of the X networks (ScoreNetwork_X.py:40-44, :84-87): xs [n][fdim] -> 2f -> 2f -> F, ELU, mask_x
DEVINL void final_node_mlp(const FNet& net, const float* xs, int XSS, int n, const float* fl, float* hA, float* hB,
                           float* out_g, int F) {
  const int H = 2 * net.fdim, HP = (H + 3) & ~3;
  mm_rows<ACT_ELU>(xs, 0, XSS, net.fdim, net.fw[0], 0, H, net.fb[0], 0, 1, n, 0, HP, H, hA, 0, HP, HP, nullptr, 0, 0, nullptr);
  __syncthreads();
  mm_rows<ACT_ELU>(hA, 0, HP, H, net.fw[1], 0, H, net.fb[1], 0, 1, n, 0, HP, H, hB, 0, HP, HP, nullptr, 0, 0, nullptr);
  __syncthreads();
  // the F-wide result goes straight to global memory (row stride F)
  mm_rows<ACT_NONE>(hB, 0, HP, H, net.fw[2], 0, net.fout, net.fb[2], 0, 1, n, 0, net.fout, net.fout, out_g, 0, F, net.fout,
                    nullptr, 0, 0, fl);
  __syncthreads();
}

// ---------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fused_layers_kernel(const __grid_constant__ FusedArgs A) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x;
  const int n = A.neff[b];
  if (n <= 0) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int N = A.N, F = A.F;
  const int P = n * (n - 1) / 2;
  const FSmem& S = A.sl;
  const int PM = S.PM, LD = S.LD, XS = S.XS, QKS = S.QKS, NBP = S.NBP, XSS = S.XSS;
  float* fl = sm + S.fl;
  float* dinv = sm + S.dinv;
  unsigned* pix = reinterpret_cast<unsigned*>(sm + S.pix);
  float* x0 = sm + S.x0;
  float* xb = sm + S.xb;
  float* xc = sm + S.xc;
  float* xs = sm + S.xs;
  float* h1 = sm + S.h1;
  float* wts = sm + S.wts;
  float* P0 = sm + S.p0;
  float* PA = sm + S.pa;
  float* R1 = sm + S.r1;           // A.x products, then attention maps
  float* QK = sm + S.qk;           // Q|K rows; [QK | R2] also hosts the dense normalised adjacency
  float* R2 = sm + S.r2;           // V rows
  float* dense = QK;
  const int64_t gp0 = A.poff ? A.poff[b] : 0;

  // ---- stage the graph
  for (int i = tid; i < N; i += nt) fl[i] = A.flags[(int64_t)b * N + i];
  for (int i = tid; i < n; i += nt) {
    const int rs = i * (2 * n - i - 1) / 2;
    for (int j = i + 1; j < n; ++j) pix[rs + j - i - 1] = (unsigned)((i << 8) | j);
  }
  for (int idx = tid; idx < n * XS; idx += nt) {
    const int i = idx / XS, k = idx - i * XS;
    x0[idx] = k < F ? A.x_in[((int64_t)b * N + i) * F + k] : 0.f;
  }
  {
    const float* Ag = A.adj + (int64_t)b * N * N;
    for (int idx = tid; idx < n * n; idx += nt) {
      const int i = idx / n, j = idx - i * n;
      dense[i * LD + j] = Ag[i * N + j];
    }
  }
  __syncthreads();
  // ---- power stack (graph_utils.py:133-142), c_init <= 2: plane 0 = A, plane 1 = A @ A
  int c_init = 1;
  if (A.has_a) c_init = A.a.c_init;
  if (A.has_x && A.x.type == GDSS_NET_X_GMH && A.x.c_init > c_init) c_init = A.x.c_init;
  for (int p = tid; p < P; p += nt) {
    const unsigned ij = pix[p];
    const int i = ij >> 8, j = ij & 255;
    const float a = dense[i * LD + j];
    P0[p] = a;
    float a2 = 0.f;
    if (c_init > 1) {
      for (int m = 0; m < n; ++m) a2 = fmaf(dense[i * LD + m], dense[m * LD + j], a2);
      P0[PM + p] = a2;
    }
    if (A.has_a) {
      A.stackT[gp0 + p] = a;
      if (A.a.c_init > 1) A.stackT[A.pstride + gp0 + p] = a2;
    }
  }
  __syncthreads();

  // ---- the two networks
  for (int which = 0; which < 2; ++which) {
    const bool isA = which == 0;
    if (isA ? !A.has_a : !A.has_x) continue;
    const FNet& net = isA ? A.a : A.x;
    if (!isA) {
      // node feature stack [x, x_1 .. x_depth] (ScoreNetwork_X.py:40, :84), padded columns zero
      for (int idx = tid; idx < n * XSS; idx += nt) {
        const int i = idx / XSS, k = idx - i * XSS;
        xs[idx] = k < F ? x0[i * XS + k] : 0.f;
      }
    }
    if (net.type == GDSS_NET_X) {
      // ScoreNetworkX (ScoreNetwork_X.py:32-46): x_l = tanh(DenseGCNConv_l(x_{l-1}, adj)), raw adjacency only
      build_norm_adj(P0, PM, 1, n, LD, NBP, dense, dinv);
      const float* xin = x0;
      for (int l = 0; l < net.depth; ++l) {
        const int f = l == 0 ? F : net.nhid, fp = (f + 3) & ~3;
        adj_times_x(dense, LD, xin, XS, 1, n, fp, R1);
        __syncthreads();
        float* xo = (l & 1) ? xc : xb;
        mm_rows<ACT_TANH>(R1, 0, fp, f, net.gw[l], 0, net.nhid, net.gb[l], 0, 1, n, 0, net.nhid, net.nhid, xo, 0, XS, net.nhid,
                          nullptr, 0, 0, nullptr);
        __syncthreads();
        for (int idx = tid; idx < n * net.nhid; idx += nt) {
          const int i = idx / net.nhid, o = idx - i * net.nhid;
          xs[i * XSS + F + l * net.nhid + o] = xo[i * XS + o];
        }
        xin = xo;
        __syncthreads();
      }
      final_node_mlp(net, xs, XSS, n, fl, PA, PA + n * ((2 * net.fdim + 3) & ~3), A.raw_x + (int64_t)b * N * F, F);
      continue;
    }
    // AttentionLayer stack (ScoreNetwork_A.py:131-141, ScoreNetwork_X.py:75-84)
    const float* pin = P0;
    const float* xin = x0;
    int cbase = net.c_init;
    const float scale = 1.0f / sqrtf((float)net.nhid);
    for (int l = 0; l < net.depth; ++l) {
      const FLayer& L = net.L[l];
      const int C = L.c_in, f = L.f_in, fp = (f + 3) & ~3, attn = L.attn, h = net.nhid;
      if (L.need_edge) stage_edge_weights(L, wts);
      build_norm_adj(pin, PM, C, n, LD, NBP, dense, dinv);
      adj_times_x(dense, LD, xin, XS, C, n, fp, R1);
      __syncthreads();
      // Q | K | V = (A^ x) W + b for every channel (attention.py:32-34); V rows are stored [i][c][h]
      mm_rows<ACT_NONE>(R1, n * fp, fp, f, L.wqkv, f * L.qkvw, L.qkvw, L.bqkv, L.qkvw, C, n, L.need_edge ? 0 : 2 * attn,
                        L.need_node ? L.qkvw : 2 * attn, L.qkvw, QK, n * QKS, QKS, 2 * attn, R2, h, C * h, nullptr);
      __syncthreads();
      float* xo = (l & 1) ? xc : xb;
      if (L.need_node) {
        // x_out = tanh(mask_x(multi_channel(cat_c V_c)))  (attention.py:106-107), tanh'd twice in X_GMH (:81)
        mm_rows<ACT_ELU>(R2, 0, C * h, C * h, L.nw0, 0, 16, L.nb0, 0, 1, n, 0, 16, 16, h1, 0, 16, 16, nullptr, 0, 0, nullptr);
      }
      if (L.need_edge) attention_maps(QK, QKS, pix, C, n, P, PM, attn, net.heads, scale, R1);
      __syncthreads();
      if (L.need_node) {
        if (isA)
          mm_rows<ACT_TANH>(h1, 0, 16, 16, L.nw1, 0, h, L.nb1, 0, 1, n, 0, h, h, xo, 0, XS, h, nullptr, 0, 0, fl);
        else
          mm_rows<ACT_TANH2>(h1, 0, 16, 16, L.nw1, 0, h, L.nb1, 0, 1, n, 0, h, h, xo, 0, XS, h, nullptr, 0, 0, fl);
      }
      if (L.need_edge) {
        const bool keep = l + 1 < net.depth;                    // a later layer reads the planes
        float* gout = isA ? A.stackT + (int64_t)cbase * A.pstride + gp0 : nullptr;
        if (L.nlin == 3) edge_mlp<3>(R1, pin, PM, pix, fl, wts, C, P, L.c_out, keep ? PA : nullptr, gout, A.pstride);
        else edge_mlp<2>(R1, pin, PM, pix, fl, wts, C, P, L.c_out, keep ? PA : nullptr, gout, A.pstride);
      }
      __syncthreads();
      if (L.need_node) {
        if (!isA) {
          for (int idx = tid; idx < n * h; idx += nt) {
            const int i = idx / h, o = idx - i * h;
            xs[i * XSS + F + l * h + o] = xo[i * XS + o];
          }
        }
        xin = xo;
      }
      pin = PA;
      cbase += L.c_out;
      __syncthreads();
    }
    if (!isA) final_node_mlp(net, xs, XSS, n, fl, PA, PA + n * ((2 * net.fdim + 3) & ~3), A.raw_x + (int64_t)b * N * F, F);
  }
}

// ---------------------------------------------------------------------------------------
// final_edge_tri_kernel: ScoreNetworkA.final (ScoreNetwork_A.py:141-148) over 128-pixel tiles of the whole
// batch, FP32 FFMA (register-blocked 8 pixels x OPT hidden units per thread).  Reads stackT[c][gp].
// ---------------------------------------------------------------------------------------
struct FinalTriArgs {
  const float* stackT; int64_t pstride;
  const int* gmap; const int* meta;
  const float* w1; const float* b1; const float* w2; const float* b2; const float* w3; const float* b3;
  float* out; const float* flags;
  int N, fdim, H;
};

template <int OPT>
__global__ void __launch_bounds__(256) final_edge_tri_kernel(FinalTriArgs a) {
  extern __shared__ __align__(16) float sm[];
  constexpr int HP = 16 * OPT;
  const int tile = blockIdx.x;
  if (tile >= a.meta[1]) return;
  const int64_t gp0 = (int64_t)tile * 128;
  const int K1 = a.fdim, H = a.H, N = a.N;
  const int sizeA = max(K1 * 128 + K1 * HP, H * HP);
  float* Xs = sm;                   // [K1][128]
  float* W1s = Xs + K1 * 128;       // [K1][HP]
  float* W2s = sm;                  // [H][HP]   (aliases Xs/W1s after layer 1)
  float* H1s = sm + sizeA;          // [HP][128]
  float* b1s = H1s + HP * 128;      // [HP]
  float* b2s = b1s + HP;
  float* w3s = b2s + HP;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < HP; idx += 256) {
    b1s[idx] = idx < H ? a.b1[idx] : 0.f;
    b2s[idx] = idx < H ? a.b2[idx] : 0.f;
    w3s[idx] = idx < H ? a.w3[idx] : 0.f;
  }
  for (int idx = tid; idx < K1 * HP; idx += 256) {
    int k = idx / HP, o = idx - k * HP;
    W1s[idx] = o < H ? a.w1[k * H + o] : 0.f;
  }
  for (int idx = tid; idx < K1 * 128; idx += 256) {
    int k = idx >> 7, m = idx & 127;
    Xs[idx] = a.stackT[(int64_t)k * a.pstride + gp0 + m];       // padded planes: always readable
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][OPT];
#pragma unroll
  for (int q = 0; q < OPT; ++q) {
    float bb = b1s[ty * OPT + q];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r][q] = bb;
  }
  for (int k = 0; k < K1; ++k) {
    const float4 xa = *reinterpret_cast<const float4*>(Xs + k * 128 + 4 * tx);
    const float4 xb = *reinterpret_cast<const float4*>(Xs + k * 128 + 64 + 4 * tx);
    const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
    float wv[OPT];
#pragma unroll
    for (int q = 0; q < OPT; ++q) wv[q] = W1s[k * HP + ty * OPT + q];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int q = 0; q < OPT; ++q) acc[r][q] = fmaf(xv[r], wv[q], acc[r][q]);
  }
#pragma unroll
  for (int q = 0; q < OPT; ++q) {
    float4 ha = make_float4(elu_f(acc[0][q]), elu_f(acc[1][q]), elu_f(acc[2][q]), elu_f(acc[3][q]));
    float4 hb = make_float4(elu_f(acc[4][q]), elu_f(acc[5][q]), elu_f(acc[6][q]), elu_f(acc[7][q]));
    *reinterpret_cast<float4*>(H1s + (ty * OPT + q) * 128 + 4 * tx) = ha;
    *reinterpret_cast<float4*>(H1s + (ty * OPT + q) * 128 + 64 + 4 * tx) = hb;
  }
  __syncthreads();
  for (int idx = tid; idx < H * HP; idx += 256) {
    int k = idx / HP, o = idx - k * HP;
    W2s[idx] = o < H ? a.w2[k * H + o] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < OPT; ++q) {
    float bb = b2s[ty * OPT + q];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r][q] = bb;
  }
  for (int k = 0; k < H; ++k) {
    const float4 xa = *reinterpret_cast<const float4*>(H1s + k * 128 + 4 * tx);
    const float4 xb = *reinterpret_cast<const float4*>(H1s + k * 128 + 64 + 4 * tx);
    const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
    float wv[OPT];
#pragma unroll
    for (int q = 0; q < OPT; ++q) wv[q] = W2s[k * HP + ty * OPT + q];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int q = 0; q < OPT; ++q) acc[r][q] = fmaf(xv[r], wv[q], acc[r][q]);
  }
  float part[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) part[r] = 0.f;
#pragma unroll
  for (int q = 0; q < OPT; ++q) {
    float w = w3s[ty * OPT + q];
#pragma unroll
    for (int r = 0; r < 8; ++r) part[r] = fmaf(elu_f(acc[r][q]), w, part[r]);
  }
  __syncthreads();
  float* red = H1s;
  *reinterpret_cast<float4*>(red + ty * 128 + 4 * tx) = make_float4(part[0], part[1], part[2], part[3]);
  *reinterpret_cast<float4*>(red + ty * 128 + 64 + 4 * tx) = make_float4(part[4], part[5], part[6], part[7]);
  __syncthreads();
  if (tid < 128) {
    const int v = a.gmap[gp0 + tid];
    if (v >= 0) {
      float s = a.b3[0];
#pragma unroll
      for (int g = 0; g < 16; ++g) s += red[g * 128 + tid];
      const int b = v >> 12, i = (v >> 6) & 63, j = v & 63;
      const float* fl = a.flags + (int64_t)b * N;
      s *= fl[i] * fl[j];
      float* O = a.out + (int64_t)b * N * N;
      O[i * N + j] = s;
      O[j * N + i] = s;
    }
  }
}

typedef void (*final_tri_fn)(FinalTriArgs);
static final_tri_fn pick_final_tri(int H, int* hp) {
  int opt = (H + 15) / 16;
  *hp = 16 * opt;
  switch (opt) {
    case 1: return final_edge_tri_kernel<1>;
    case 2: return final_edge_tri_kernel<2>;
    case 3: return final_edge_tri_kernel<3>;
    case 4: return final_edge_tri_kernel<4>;
    case 5: return final_edge_tri_kernel<5>;
    case 6: return final_edge_tri_kernel<6>;
    case 7: return final_edge_tri_kernel<7>;
    case 8: return final_edge_tri_kernel<8>;
  }
  return nullptr;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static bool net_fusable(const NetPlan& p) {
  if (p.type == GDSS_NET_X) return p.depth <= FUSED_MAX_LAYERS && p.nhid % 4 == 0 && p.nhid <= 64;
  if (p.depth > FUSED_MAX_LAYERS || p.c_init > 2 || p.heads != 4) return false;
  if (p.nhid % 16 || p.adim % 16 || p.nhid > 64 || p.adim > 64) return false;
  if (p.nlin < 2 || p.nlin > 3) return false;
  for (int l = 0; l < p.depth; ++l) {
    const LayerPlan& L = p.layers[l];
    if (L.hid != 16 || L.c_in > 8 || L.c_out > 8) return false;
  }
  return true;
}

static void fill_net(const NetPlan& p, const float* blob, FNet* f) {
  memset(f, 0, sizeof(*f));
  f->type = p.type; f->depth = p.depth; f->F = p.F; f->nhid = p.nhid; f->heads = p.heads;
  f->fdim = p.fdim; f->fout = p.fout; f->c_init = p.c_init;
  for (int l = 0; l < p.depth && l < FUSED_MAX_LAYERS; ++l) {
    if (p.type == GDSS_NET_X) {
      f->gw[l] = blob + p.gw[l];
      f->gb[l] = blob + p.gb[l];
      continue;
    }
    const LayerPlan& L = p.layers[l];
    FLayer& o = f->L[l];
    const bool last = l == p.depth - 1;
    o.c_in = L.c_in; o.c_out = L.c_out; o.f_in = L.f_in; o.attn = L.attn; o.nlin = L.nlin; o.qkvw = L.qkvw;
    o.need_edge = (p.type == GDSS_NET_A || !last) ? 1 : 0;   // X_GMH: the last layer's adj_out is never read
    o.need_node = (p.type != GDSS_NET_A || !last) ? 1 : 0;   // A: the last layer's x_out is never read
    o.wqkv = blob + L.wqkv; o.bqkv = blob + L.bqkv;
    for (int i = 0; i < L.nlin && i < 3; ++i) { o.ew[i] = blob + L.ew[i]; o.eb[i] = blob + L.eb[i]; }
    o.nw0 = blob + L.nw[0]; o.nb0 = blob + L.nb[0]; o.nw1 = blob + L.nw[1]; o.nb1 = blob + L.nb[1];
  }
  for (int i = 0; i < 3; ++i) { f->fw[i] = blob + p.fw[i]; f->fb[i] = blob + p.fb[i]; }
}

// shared-memory plan for graphs of at most NB nodes; returns bytes (or -1 when the regions cannot host the
// dense adjacency alias)
static int64_t plan_smem(const gdss_ctx* ctx, int NB, FSmem* s) {
  const NetPlan* nets[2] = {ctx->has_a ? &ctx->pa : nullptr, ctx->has_x ? &ctx->px : nullptr};
  int Cmax = 1, hmax = 4, amax = 0, fmax = (ctx->F + 3) & ~3, fdx = 4, Hx = 0, cinit = 1;
  for (int k = 0; k < 2; ++k) {
    const NetPlan* p = nets[k];
    if (!p) continue;
    hmax = hmax > p->nhid ? hmax : p->nhid;
    if (p->type != GDSS_NET_A) { fdx = (p->fdim + 3) & ~3; Hx = (2 * p->fdim + 3) & ~3; }
    if (p->type == GDSS_NET_X) continue;
    cinit = cinit > p->c_init ? cinit : p->c_init;
    for (int l = 0; l < p->depth; ++l) {
      const LayerPlan& L = p->layers[l];
      Cmax = Cmax > L.c_in ? Cmax : L.c_in;
      Cmax = Cmax > L.c_out ? Cmax : L.c_out;
      amax = amax > L.attn ? amax : L.attn;
    }
  }
  if (fmax < hmax) fmax = hmax;
  const int NBP = (NB + 3) & ~3;
  const int PM = ((NB * (NB - 1) / 2) + 3) & ~3;
  s->NBP = NBP; s->PM = PM > 4 ? PM : 4;
  s->LD = NB | 1;
  s->XS = fmax;
  s->QKS = 2 * amax + 4;
  s->XSS = fdx;
  int cur = 0;
  auto take = [&](int nfl) { int o = cur; cur += (nfl + 3) & ~3; return o; };
  s->fl = take(NBP);
  s->dinv = take(Cmax * NBP);
  s->pix = take(s->PM);
  s->x0 = take(NB * s->XS);
  s->xb = take(NB * s->XS);
  s->xc = take(NB * s->XS);
  s->xs = take(NB * s->XSS);
  s->h1 = take(NB * 16);
  s->wts = take(2 * Cmax * 16 + 16 + 256 + 16 + 128 + 8);
  s->p0 = take(cinit * s->PM);
  int pa = Cmax * s->PM;
  if (pa < 2 * NB * Hx) pa = 2 * NB * Hx;            // hosts the two hidden buffers of the final node MLP
  s->pa = take(pa);
  int r1 = Cmax * NB * fmax;                         // A.x products [C][n][fp]
  if (r1 < Cmax * s->PM) r1 = Cmax * s->PM;          // attention maps [C][P]
  s->r1 = take(r1);
  const int dense = Cmax * NB * s->LD;
  int qk = Cmax * NB * s->QKS;
  int r2 = Cmax * NB * hmax;
  if (qk + r2 < dense) r2 = dense - qk;
  s->qk = take(qk);
  s->r2 = take(r2);
  s->total = cur;
  return (int64_t)cur * 4;
}

bool fused_supported(const gdss_ctx* ctx) {
  if (ctx->N > 64 || ctx->N < 2) return false;
  if (ctx->has_a && !net_fusable(ctx->pa)) return false;
  if (ctx->has_x && !net_fusable(ctx->px)) return false;
  FSmem s;
  const int64_t bytes = plan_smem(ctx, ctx->N, &s);
  return bytes > 0 && bytes <= ctx->smem_optin;
}

int launch_tri_setup(const gdss_ctx* ctx, const int* neff, int B, char* ws, const Workspace& w, cudaStream_t st) {
  int* poff = reinterpret_cast<int*>(ws + w.off[R_POFF]);
  int* meta = reinterpret_cast<int*>(ws + w.off[R_META]);
  int* gmap = reinterpret_cast<int*>(ws + w.off[R_GMAP]);
  tri_scan_kernel<<<1, 1024, 0, st>>>(neff, B, poff, meta, gmap);
  launched(K_MISC, st);
  LAUNCH_CHECK();
  tri_map_kernel<<<B, 64, 0, st>>>(neff, poff, gmap);
  launched(K_MISC, st);
  LAUNCH_CHECK();
  return 0;
}

int launch_eval_fused(const gdss_ctx* ctx, const float* x, const float* adj, const float* flags, const int* neff,
                      float* score_x, float* score_adj, int B, char* ws, const Workspace& w, cudaStream_t st) {
  const int N = ctx->N;
  const bool want_a = score_adj && ctx->has_a, want_x = score_x && ctx->has_x;
  if (!want_a && !want_x) return 0;
  FusedArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.has_a = want_a ? 1 : 0;
  fa.has_x = want_x ? 1 : 0;
  if (want_a) fill_net(ctx->pa, ctx->blob_a, &fa.a);
  if (want_x) fill_net(ctx->px, ctx->blob_x, &fa.x);
  fa.x_in = x; fa.adj = adj; fa.flags = flags; fa.neff = neff;
  fa.raw_x = score_x;
  fa.stackT = reinterpret_cast<float*>(ws + w.off[R_STACKT]);
  fa.poff = reinterpret_cast<const int*>(ws + w.off[R_POFF]);
  fa.pstride = w.pstride;
  fa.N = N; fa.F = ctx->F;
  const int64_t smem = plan_smem(ctx, N, &fa.sl);
  if (smem <= 0 || smem > ctx->smem_optin) return GDSS_E_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(fused_layers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int threads = N <= 12 ? 64 : (N <= 24 ? 128 : 256);
  fused_layers_kernel<<<B, threads, smem, st>>>(fa);
  launched(K_FUSED, st);
  LAUNCH_CHECK();
  if (want_a) {
    const NetPlan& p = ctx->pa;
    FinalTriArgs ta;
    ta.stackT = fa.stackT; ta.pstride = w.pstride;
    ta.gmap = reinterpret_cast<const int*>(ws + w.off[R_GMAP]);
    ta.meta = reinterpret_cast<const int*>(ws + w.off[R_META]);
    ta.w1 = ctx->blob_a + p.fw[0]; ta.b1 = ctx->blob_a + p.fb[0]; ta.w2 = ctx->blob_a + p.fw[1]; ta.b2 = ctx->blob_a + p.fb[1];
    ta.w3 = ctx->blob_a + p.fw[2]; ta.b3 = ctx->blob_a + p.fb[2];
    ta.out = score_adj; ta.flags = flags; ta.N = N; ta.fdim = p.fdim; ta.H = 2 * p.fdim;
    int HP;
    final_tri_fn fn = pick_final_tri(ta.H, &HP);
    if (!fn) return GDSS_E_UNSUPPORTED;
    int64_t sizeA = (int64_t)p.fdim * 128 + (int64_t)p.fdim * HP;
    if (sizeA < (int64_t)ta.H * HP) sizeA = (int64_t)ta.H * HP;
    const size_t fsm = sizeof(float) * (sizeA + (size_t)HP * 128 + 3 * HP);
    if ((int)fsm > ctx->smem_optin) return GDSS_E_UNSUPPORTED;
    if (fsm > 48 * 1024) {
      e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);
      if (e != cudaSuccess) return (int)e;
    }
    const int64_t max_tiles = ((int64_t)B * (N * (N - 1) / 2) + 127) / 128;
    if (max_tiles > 0) {
      fn<<<(unsigned)max_tiles, 256, fsm, st>>>(ta);
      launched(K_FINAL_EDGE, st);
      LAUNCH_CHECK();
    }
  }
  return 0;
}

}  // namespace gdss
