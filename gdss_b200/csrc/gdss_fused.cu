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
#include <stdlib.h>
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
  return fmaf(-2.f, rcp_ftz(ex2_ftz(v * 2.8853900817779268f) + 1.f), 1.f);
}
// 1 / (exp(2 z) + 1) with the 2 log2(e) factor already folded into the argument: tanh(z) = 1 - 2 * sig2(z * 2 log2 e)
DEVINL float sig2(float a) { return rcp_ftz(ex2_ftz(a) + 1.f); }

// ---------------------------------------------------------------------------------------
// pixel bookkeeping (depends on the flags only: once per sampler call)
// ---------------------------------------------------------------------------------------
// meta: [0] total pixels, [1] 128-pixel tiles, [2..4] graphs per size bucket, [5..7] first slot of each bucket in order[]
// order[]: graph ids grouped by size bucket and, inside a bucket, sorted by node count in DESCENDING order (counting sort;
// the longest CTAs of a launch start first, so the last wave of a launch is made of its shortest graphs).  The position
// of a graph among those of equal size is arbitrary; results do not depend on it.  wide: gmap fields of 9 bits (general
// path, N <= 511) instead of 6.
#define TRI_BINS 66
__global__ void __launch_bounds__(1024) tri_scan_kernel(const int* __restrict__ neff, int B, int nb0, int nb1,
                                                        int* __restrict__ poff, int* __restrict__ meta,
                                                        int* __restrict__ gmap, int* __restrict__ order,
                                                        int* __restrict__ noff, int* __restrict__ nmap) {
  __shared__ int part[1024];
  __shared__ int npart[1024];          // node counts: the compact row index of node (b, i) is noff[b] + i
  __shared__ int hist[TRI_BINS], cursor[TRI_BINS];
  __shared__ int bstart[3], bcount[3];
  const int t = threadIdx.x;
  const int chunk = (B + 1023) / 1024;
  const int lo = min(B, t * chunk), hi = min(B, lo + chunk);
  if (t < TRI_BINS) hist[t] = 0;
  __syncthreads();
  int s = 0, sn = 0;
  for (int b = lo; b < hi; ++b) {
    const int n = neff[b];
    s += n * (n - 1) / 2;
    sn += max(n, 0);
    atomicAdd(&hist[min(max(n, 0), TRI_BINS - 1)], 1);
  }
  // exclusive prefix of the pixel / node counts over the threads' chunks
  part[t] = s;
  npart[t] = sn;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    int v = t >= off ? part[t - off] : 0;
    int vn = t >= off ? npart[t - off] : 0;
    __syncthreads();
    part[t] += v;
    npart[t] += vn;
    __syncthreads();
  }
  int base = part[t] - s;
  int nbase = npart[t] - sn;
  const int tot = part[1023], ntot = npart[1023];
  if (t == 0) {
    int cur = 0;
    for (int k = 0; k < 3; ++k) {
      const int blo = k == 0 ? 0 : (k == 1 ? nb0 + 1 : nb1 + 1);
      const int bhi = min(TRI_BINS - 1, k == 0 ? nb0 : (k == 1 ? nb1 : TRI_BINS - 1));
      bstart[k] = cur;
      for (int n = bhi; n >= blo; --n) {
        cursor[n] = cur;
        cur += hist[n];
      }
      bcount[k] = cur - bstart[k];
    }
  }
  __syncthreads();
  for (int b = lo; b < hi; ++b) {
    poff[b] = base;
    if (noff) noff[b] = nbase;
    const int n = neff[b];
    base += n * (n - 1) / 2;
    nbase += max(n, 0);
    order[atomicAdd(&cursor[min(max(n, 0), TRI_BINS - 1)], 1)] = b;
  }
  if (t == 1023) {
    poff[B] = tot;
    meta[0] = tot;
    meta[1] = (tot + 127) / 128;
    meta[8] = ntot;                        // valid nodes of the batch / 128-row tiles of the compact node stack
    meta[9] = (ntot + 127) / 128;
    if (noff) noff[B] = ntot;
    if (nmap)
      for (int k = ntot; k < ((ntot + 127) / 128) * 128; ++k) nmap[k] = -1;
    meta[16] = 0;                          // error flag of the tensor-core kernel (lost mbarrier arrival)
    for (int q = 0; q < 3; ++q) {
      meta[2 + q] = bcount[q];
      meta[5 + q] = bstart[q];
    }
    for (int k = tot; k < ((tot + 127) / 128) * 128; ++k) gmap[k] = -1;
  }
}

// gmap[gp] = b << 2 fb | i << fb | j  (fb = 6 bits for the per-graph path, 9 for the general path)
__global__ void __launch_bounds__(64) tri_map_kernel(const int* __restrict__ neff, const int* __restrict__ poff,
                                                     int* __restrict__ gmap, int fb, const int* __restrict__ noff,
                                                     int* __restrict__ nmap) {
  const int b = blockIdx.x, n = neff[b], base = poff[b];
  if (nmap)                                                // compact node rows: nmap[noff[b] + i] = b << 6 | i
    for (int i = threadIdx.x; i < n; i += blockDim.x) nmap[noff[b] + i] = (b << 6) | i;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int rs = i * (2 * n - i - 1) / 2;
    for (int j = i + 1; j < n; ++j) gmap[base + rs + j - i - 1] = (b << (2 * fb)) | (i << fb) | j;
  }
}

// ---------------------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------------------
#define FUSED_MAX_LAYERS 8
#define RMAX 8                 // rows per register-blocked item

// The weights of one AttentionLayer are two contiguous runs of the blob (gdss_pack.cu lays a layer out as
// wqkv, bqkv | mlp linears, multi_channel linears); each run is copied to shared memory by one bulk async copy
// (TMA, cp.async.bulk) one phase group ahead of its use.  *_o = float offsets inside the run.
struct FLayer {
  int c_in, c_out, f_in, f_in_p, attn, nlin, qkvw, need_edge, need_node;
  // dead sub-network elimination (gdss_pack.cu analyze_liveness): channels whose Q/K chain + map (qk_live) or V
  // convolution (v_live) can reach an output; act[] = channels with any live chain, qkc[] = those with a live map;
  // edge_const: the per-edge MLP's output is the constant cst[] (times the flag mask)
  unsigned qk_live, v_live;
  int nact, nqk, edge_const;
  unsigned long long vk_live;          // k-steps (8 inputs each) of the node MLP's first linear that belong to live V blocks
  signed char act[8], qkc[8];
  float cst[8];
  const float* blk_q; int q_floats, bq_o;                 // [c_in][f_in_p][qkvw] then biases [c_in][qkvw]
  const float* blk_m; int m_floats, ew_o[3], eb_o[3], nw0_o, nb0_o, nw1_o, nb1_o;
};

struct FNet {
  int type, depth, F, nhid, heads, fdim, fout, c_init, Hp, foutp;
  FLayer L[FUSED_MAX_LAYERS];
  const float* gw[FUSED_MAX_LAYERS];
  const float* gb[FUSED_MAX_LAYERS];
  const float *fw[3], *fb[3];
};

// float offsets into dynamic shared memory
struct FSmem {
  int fl, dinv, pix, x0, xb, xs, h1, wq, wm, p0, pa, r1, qk, r2, tab, total;
  int LD, XS, QKS, PM, NBP, XSS, RS;     // RS: per-channel stride of R1 (A.x rows / attention maps of one channel)
  int TC;                                // columns of the triangle-index table (8 * ceil(NB / 8)); rows = 16 * ceil(NB / 16)
};

struct FusedArgs {
  FNet a, x;
  int has_a, has_x;
  const float *x_in, *adj, *flags;
  const int* neff;
  float* raw_x;
  float* stackT;
  float* xst;              // X network: node-feature stack [fdim][nstride] in global memory (final MLP runs in the batched
  int64_t nstride;         // tensor-core kernel) or null (final MLP inside this kernel); row = b * N + i
  const int* poff;
  const int* noff;         // compact row offset of a graph's nodes in the X network's node stack
  const int* order;        // graph ids sorted by size bucket
  const int* meta;
  int bucket;
  int mb1;                  // every graph of the launch has at most 16 nodes (one mma row tile)
  int use_mma;              // warp-level tensor-core (mma.sync 3xTF32) variants of the per-graph contractions
  int use_tc_edge;          // per-edge MLPs of the AttentionLayers on tcgen05 (CTAs of >= 256 threads, hidden width 16)
  int tce_slots;            // tiles in flight per round (1 .. EDGE_SLOTS)
  unsigned long long* prof;   // optional [16] per-phase clock cycles (thread 0 of every CTA, summed), else null
  int64_t pstride;
  int N, F;
  FSmem sl;
};

// ---------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------
// exact x / d for 0 <= x < 2^16, 1 <= d < 2^16 (one IMAD.HI instead of an integer division)
struct FastDiv {
  unsigned m;
  int d;
  DEVINL explicit FastDiv(int dd) : m(dd > 1 ? (0xFFFFFFFFu / (unsigned)dd + 1u) : 0u), d(dd) {}
  DEVINL int div(int x) const { return d > 1 ? (int)__umulhi((unsigned)x, m) : x; }
};

enum { ACT_NONE = 0, ACT_ELU = 1, ACT_TANH = 2, ACT_TANH2 = 3 };

// Blackwell packed FP32: one FFMA2 issues two independent IEEE fp32 FMAs (same results as two scalar FMAs)
DEVINL float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
DEVINL float2 dup2(float v) { return make_float2(v, v); }
DEVINL float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
DEVINL float2 hi2(const float4& v) { return make_float2(v.z, v.w); }

DEVINL uint32_t sm_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one elected thread: arm the mbarrier with the byte count and start the bulk global -> shared copy (TMA)
DEVINL void tma_fetch(float* dst, const float* src, int floats, uint64_t* mbar) {
  const uint32_t bytes = (uint32_t)floats * 4u;
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // earlier generic reads of dst are done (after the CTA barrier)
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(sm_addr(mbar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(sm_addr(dst)),
               "l"(src), "r"(bytes), "r"(sm_addr(mbar))
               : "memory");
}

DEVINL void mbar_wait_parity(uint64_t* mbar, uint32_t parity) {
  const uint32_t a = sm_addr(mbar);
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// mm_rows: out[c][i][o] = act( (bias[c][o] + sum_k in[c][i][k] * W[c][k][o]) * rowscale[i] )
// Register-blocked: one item = 4 consecutive output columns x up to RMAX rows; the 4x4 weight block of a
// k-chunk is loaded once (global; L1/L2-resident, shared by every CTA) and reused over the rows, the input
// rows come from shared memory as float4.  `in` rows must be readable (and finite) up to K rounded up to 4.
// Output columns < split go to out0, the others to out1 (column index minus split).  One out-of-line
// instance (instruction-cache footprint), the activation is a run-time switch in the epilogue.
// ---------------------------------------------------------------------------------------
struct MM {
  const float* in; int in_cs, ldin, K;       // K rounded up to 4 by the caller: W rows [K_true, K) are zero in the blob
  const float* W; int w_cs, ldw;
  const float* bias; int b_cs;
  int C, n, o_lo, o_hi;
  float* out0; int o0_cs, ld0, split;
  float* out1; int o1_cs, ld1;
  const float* rowscale;
  int act, vec_out;
};

template <bool WS>          // WS: weights / biases in shared memory (staged by TMA), else global (read-only path)
DEVINL float4 ldw4(const float* p) {
  if (WS) return *reinterpret_cast<const float4*>(p);
  return __ldg(reinterpret_cast<const float4*>(p));
}

template <bool WS>
DEVINL void mm_rows_t(const MM& m) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n = m.n, K4 = m.K, ldw = m.ldw, ldin = m.ldin;
  const int ogn = (m.o_hi - m.o_lo + 3) >> 2;
  const int pairs = m.C * ogn;
  if (pairs <= 0 || n <= 0) return;
  // rows are dealt round-robin to IP row-partitions; pick the rows-per-item that minimises
  // (rounds of the CTA) x (cost of one item); nt is a power of two
  int IP = n, best = 1 << 30;
  const int sh = 31 - __clz(nt);
#pragma unroll
  for (int rows = 1; rows <= RMAX; ++rows) {
    const int ip = (n + rows - 1) / rows;
    const int rounds = (pairs * ip + nt - 1) >> sh;
    const int cost = rounds * (rows * 17 + 6);
    if (cost < best) { best = cost; IP = ip; }
  }
  const FastDiv dpairs(pairs), dogn(ogn);
  for (int item = tid; item < pairs * IP; item += nt) {
    const int ip = dpairs.div(item), pr = item - ip * pairs;
    const int c = dogn.div(pr), og = pr - c * ogn;
    const int o = m.o_lo + 4 * og;
    const float* Wc = m.W + (int64_t)c * m.w_cs + o;
    const float* inc = m.in + c * m.in_cs + ip * ldin;
    const int rcnt = (n - ip + IP - 1) / IP;       // rows of this item: ip, ip + IP, ...
    const int rstep = IP * ldin;
    float2 acc[RMAX][2];
    {
      const float4 bv = ldw4<WS>(m.bias + c * m.b_cs + o);
#pragma unroll
      for (int r = 0; r < RMAX; ++r) { acc[r][0] = lo2(bv); acc[r][1] = hi2(bv); }
    }
    float4 wn[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) wn[kk] = ldw4<WS>(Wc + (int64_t)kk * ldw);
    for (int k0 = 0; k0 < K4; k0 += 4) {
      const float4 w0 = wn[0], w1 = wn[1], w2 = wn[2], w3 = wn[3];
      if (k0 + 4 < K4) {                             // prefetch the next weight block
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) wn[kk] = ldw4<WS>(Wc + (int64_t)(k0 + 4 + kk) * ldw);
      }
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        if (r >= rcnt) break;
        const float4 a = *reinterpret_cast<const float4*>(inc + r * rstep + k0);
        acc[r][0] = fma2(dup2(a.x), lo2(w0), acc[r][0]); acc[r][1] = fma2(dup2(a.x), hi2(w0), acc[r][1]);
        acc[r][0] = fma2(dup2(a.y), lo2(w1), acc[r][0]); acc[r][1] = fma2(dup2(a.y), hi2(w1), acc[r][1]);
        acc[r][0] = fma2(dup2(a.z), lo2(w2), acc[r][0]); acc[r][1] = fma2(dup2(a.z), hi2(w2), acc[r][1]);
        acc[r][0] = fma2(dup2(a.w), lo2(w3), acc[r][0]); acc[r][1] = fma2(dup2(a.w), hi2(w3), acc[r][1]);
      }
    }
    const int act = m.act;
    const bool to0 = o < m.split;
    float* ob = to0 ? m.out0 + c * m.o0_cs + o : m.out1 + c * m.o1_cs + (o - m.split);
    const int ldo = to0 ? m.ld0 : m.ld1;
    const bool vec = m.vec_out && (o + 3 < m.o_hi);
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      if (r >= rcnt) break;
      const int i = ip + r * IP;
      const float rs = m.rowscale ? m.rowscale[i] : 1.f;
      float v[4] = {acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[q] = v[q] * rs;
        if (act == ACT_ELU) v[q] = elu_f(v[q]);
        else if (act >= ACT_TANH) {
          v[q] = tanh_fast(v[q]);
          if (act == ACT_TANH2) v[q] = tanh_fast(v[q]);
        }
      }
      if (vec) {
        *reinterpret_cast<float4*>(ob + i * ldo) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (o + q < m.o_hi) ob[i * ldo + q] = v[q];
      }
    }
  }
}

// out-of-line instance for the cold call sites (final node MLP, plain GCN layers): instruction-cache footprint
__device__ __noinline__ void mm_rows(const MM& m) { mm_rows_t<false>(m); }

// ---------------------------------------------------------------------------------------
// DenseGCNConv pieces (layers.py:49-88) straight from the strict-upper-triangle planes.
// Row i of plane c is walked in increasing j: entries (j, i) for j < i (index step n-j-2), then the
// contiguous run (i, j) for j > i.  The self loop (diagonal := 1) is added explicitly.
// ---------------------------------------------------------------------------------------
DEVINL void gcn_degrees(const float* pin, int PM, int C, int n, int NBP, float* dinv, const FastDiv& dn) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int item = tid; item < C * n; item += nt) {
    const int c = dn.div(item), i = item - c * n;
    const float* pl = pin + c * PM;
    float s = 1.f;
    int idx = i - 1;
    for (int j = 0; j < i; ++j) {
      s += pl[idx];
      idx += n - j - 2;
    }
    idx = i * (2 * n - i - 1) / 2;
    for (int j = i + 1; j < n; ++j) s += pl[idx++];
    dinv[c * NBP + i] = 1.0f / sqrtf(fmaxf(s, 1.0f));
  }
}

// ax[c][i][8g..8g+7] = d_i * sum_j A~_c(i,j) d_j x[j][8g..8g+7]
DEVINL void gcn_ax(const float* pin, int PM, const float* dinv, int NBP, const float* x, int ldx, int C, int n, int f8,
                   int ldax, float* ax, int cs, const FastDiv& dn) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const FastDiv df(f8);
  for (int item = tid; item < C * n * f8; item += nt) {
    const int ci = df.div(item), g = item - ci * f8;
    const int c = dn.div(ci), i = ci - c * n;
    const float* pl = pin + c * PM;
    const float* dv = dinv + c * NBP;
    const float* xg = x + 8 * g;
    const float di = dv[i];
    float2 acc[4];
    {
      const float4 a = *reinterpret_cast<const float4*>(xg + i * ldx);
      const float4 b = *reinterpret_cast<const float4*>(xg + i * ldx + 4);
      const float2 d2 = dup2(di);
      acc[0] = __fmul2_rn(d2, lo2(a)); acc[1] = __fmul2_rn(d2, hi2(a));
      acc[2] = __fmul2_rn(d2, lo2(b)); acc[3] = __fmul2_rn(d2, hi2(b));
    }
    // entries (j, i), j < i: index step n - j - 2; then the contiguous run (i, j), j > i
    int idx = i - 1;
    for (int j = 0; j < i; ++j) {
      const float2 w = dup2(pl[idx] * dv[j]);
      idx += n - j - 2;
      const float4 a = *reinterpret_cast<const float4*>(xg + j * ldx);
      const float4 b = *reinterpret_cast<const float4*>(xg + j * ldx + 4);
      acc[0] = fma2(w, lo2(a), acc[0]); acc[1] = fma2(w, hi2(a), acc[1]);
      acc[2] = fma2(w, lo2(b), acc[2]); acc[3] = fma2(w, hi2(b), acc[3]);
    }
    idx = i * (2 * n - i - 1) / 2;
    for (int j = i + 1; j < n; ++j) {
      const float2 w = dup2(pl[idx++] * dv[j]);
      const float4 a = *reinterpret_cast<const float4*>(xg + j * ldx);
      const float4 b = *reinterpret_cast<const float4*>(xg + j * ldx + 4);
      acc[0] = fma2(w, lo2(a), acc[0]); acc[1] = fma2(w, hi2(a), acc[1]);
      acc[2] = fma2(w, lo2(b), acc[2]); acc[3] = fma2(w, hi2(b), acc[3]);
    }
    float* o = ax + c * cs + i * ldax + 8 * g;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0].x * di, acc[0].y * di, acc[1].x * di, acc[1].y * di);
    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[2].x * di, acc[2].y * di, acc[3].x * di, acc[3].y * di);
  }
}

// first linear of `multi_channel` (attention.py:90, :106): h1[i][0..16) = elu(b + cat_c V_c[i] . W), the K
// range split over 4 adjacent lanes and reduced with shuffles (the layer has only n x 4 column groups)
DEVINL void node_l1(const float* V, int ldv, int K, const float* W, const float* b, int n, float* h1, int tid, int nt) {
  const int lane = tid & 31;
  const int total = n * 16, kq = K >> 2;
  for (int base = tid - lane; base < total; base += nt) {
    const int idx = base + lane;
    const bool valid = idx < total;
    const int id2 = valid ? idx : total - 1;
    const int i = id2 >> 4, og = (id2 >> 2) & 3, kp = id2 & 3;
    const float* vr = V + i * ldv + kp * kq;
    const float* wr = W + (int64_t)(kp * kq) * 16 + 4 * og;
    float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
    for (int k = 0; k < kq; k += 4) {
      const float4 v = *reinterpret_cast<const float4*>(vr + k);
      const float4 w0 = *reinterpret_cast<const float4*>(wr + (k + 0) * 16);
      const float4 w1 = *reinterpret_cast<const float4*>(wr + (k + 1) * 16);
      const float4 w2 = *reinterpret_cast<const float4*>(wr + (k + 2) * 16);
      const float4 w3 = *reinterpret_cast<const float4*>(wr + (k + 3) * 16);
      a0 = fma2(dup2(v.x), lo2(w0), a0); a1 = fma2(dup2(v.x), hi2(w0), a1);
      a0 = fma2(dup2(v.y), lo2(w1), a0); a1 = fma2(dup2(v.y), hi2(w1), a1);
      a0 = fma2(dup2(v.z), lo2(w2), a0); a1 = fma2(dup2(v.z), hi2(w2), a1);
      a0 = fma2(dup2(v.w), lo2(w3), a0); a1 = fma2(dup2(v.w), hi2(w3), a1);
    }
    float4 acc = make_float4(a0.x, a0.y, a1.x, a1.y);
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (valid && kp == 0) {
      const float4 bv = *reinterpret_cast<const float4*>(b + 4 * og);
      *reinterpret_cast<float4*>(h1 + i * 16 + 4 * og) =
          make_float4(elu_f(acc.x + bv.x), elu_f(acc.y + bv.y), elu_f(acc.z + bv.z), elu_f(acc.w + bv.w));
    }
  }
}

// attention maps (attention.py:35-49): M[c][p] = mean over the 4 heads and over the two orientations of
// tanh(Q_h . K_h / sqrt(out_dim)); HD = head width (attn_dim / 4)
template <int HD>
DEVINL void attention_maps(const float* qk, int QKS, const unsigned* pix, int C, int n, int P, int MS, float scale, float* M, const FastDiv& dP) {
  const int tid = threadIdx.x, nt = blockDim.x;
  constexpr int AT = 4 * HD;
  for (int item = tid; item < C * P; item += nt) {
    const int c = dP.div(item), p = item - c * P;
    const unsigned ij = pix[p];
    const int i = ij >> 8, j = ij & 255;
    const float* qi = qk + (c * n + i) * QKS;
    const float* qj = qk + (c * n + j) * QKS;
    float m = 0.f;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float2 p1 = make_float2(0.f, 0.f), p2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 a1 = *reinterpret_cast<const float4*>(qi + h * HD + d);
        const float4 b1 = *reinterpret_cast<const float4*>(qj + AT + h * HD + d);
        const float4 a2 = *reinterpret_cast<const float4*>(qj + h * HD + d);
        const float4 b2 = *reinterpret_cast<const float4*>(qi + AT + h * HD + d);
        p1 = fma2(lo2(a1), lo2(b1), p1); p1 = fma2(hi2(a1), hi2(b1), p1);
        p2 = fma2(lo2(a2), lo2(b2), p2); p2 = fma2(hi2(a2), hi2(b2), p2);
      }
      m += tanh_fast((p1.x + p1.y) * scale) + tanh_fast((p2.x + p2.y) * scale);
    }
    M[c * MS + p] = m * 0.125f;
  }
}

// per-pixel MLP of one AttentionLayer (attention.py:109-114): two lanes per pixel, 8 hidden units each.
//   in = [M_0..M_{C-1}, P_0..P_{C-1}] -> 16 -> (16) -> c_out ; out = 2 * mlp * flag_i * flag_j
// pout may alias pin (each pixel is read completely before it is written).
template <int NLIN>
DEVINL void edge_mlp(const float* M, int MS, const float* pin, int PM, const unsigned* pix, const float* fl, const float* wm,
                     const FLayer& L, int P, float* pout, float* gout, int64_t gstride, int tid, int nt) {
  const int sub = tid & 1;
  const int C = L.c_in, c_out = L.c_out;
  const float* W1 = wm + L.ew_o[0] + 8 * sub;                 // [2C][16]
  const float* b1 = wm + L.eb_o[0];
  const float* W2 = wm + L.ew_o[1] + 8 * sub;                 // [16][16] (NLIN == 3)
  const float* b2 = wm + L.eb_o[1];
  const float* W3 = wm + L.ew_o[NLIN - 1];                    // [16][c_out]
  const float* b3 = wm + L.eb_o[NLIN - 1];
  const bool act3 = 4 * sub < c_out;                          // this lane's 4 output columns exist
  const int total = 2 * P;
  for (int base = tid - (tid & 31); base < total; base += nt) {       // warp-uniform trip count
    const int idx = base + (tid & 31);
    const bool valid = idx < total;
    const int p = valid ? (idx >> 1) : (P - 1);
    float2 h2[4];
    {
      const float4 ba = *reinterpret_cast<const float4*>(b1 + 8 * sub);
      const float4 bb = *reinterpret_cast<const float4*>(b1 + 8 * sub + 4);
      h2[0] = lo2(ba); h2[1] = hi2(ba); h2[2] = lo2(bb); h2[3] = hi2(bb);
    }
    for (int k = 0; k < C; ++k) {
      const float v = M[k * MS + p];
      const float4 wa = *reinterpret_cast<const float4*>(W1 + k * 16);
      const float4 wb = *reinterpret_cast<const float4*>(W1 + k * 16 + 4);
      h2[0] = fma2(dup2(v), lo2(wa), h2[0]); h2[1] = fma2(dup2(v), hi2(wa), h2[1]);
      h2[2] = fma2(dup2(v), lo2(wb), h2[2]); h2[3] = fma2(dup2(v), hi2(wb), h2[3]);
    }
    for (int k = 0; k < C; ++k) {
      const float v = pin[k * PM + p];
      const float4 wa = *reinterpret_cast<const float4*>(W1 + (C + k) * 16);
      const float4 wb = *reinterpret_cast<const float4*>(W1 + (C + k) * 16 + 4);
      h2[0] = fma2(dup2(v), lo2(wa), h2[0]); h2[1] = fma2(dup2(v), hi2(wa), h2[1]);
      h2[2] = fma2(dup2(v), lo2(wb), h2[2]); h2[3] = fma2(dup2(v), hi2(wb), h2[3]);
    }
    __syncwarp();                     // both lanes of the pixel have read their inputs (in-place update below)
    float hf[16];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float own = elu_f((u & 1) ? h2[u >> 1].y : h2[u >> 1].x);
      const float oth = __shfl_xor_sync(0xffffffffu, own, 1);
      hf[u] = sub ? oth : own;
      hf[8 + u] = sub ? own : oth;
    }
    if (NLIN == 3) {
      const float4 ba = *reinterpret_cast<const float4*>(b2 + 8 * sub);
      const float4 bb = *reinterpret_cast<const float4*>(b2 + 8 * sub + 4);
      h2[0] = lo2(ba); h2[1] = hi2(ba); h2[2] = lo2(bb); h2[3] = hi2(bb);
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float v = hf[u];
        const float4 wa = *reinterpret_cast<const float4*>(W2 + u * 16);
        const float4 wb = *reinterpret_cast<const float4*>(W2 + u * 16 + 4);
        h2[0] = fma2(dup2(v), lo2(wa), h2[0]); h2[1] = fma2(dup2(v), hi2(wa), h2[1]);
        h2[2] = fma2(dup2(v), lo2(wb), h2[2]); h2[3] = fma2(dup2(v), hi2(wb), h2[3]);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float own = elu_f((u & 1) ? h2[u >> 1].y : h2[u >> 1].x);
        const float oth = __shfl_xor_sync(0xffffffffu, own, 1);
        hf[u] = sub ? oth : own;
        hf[8 + u] = sub ? own : oth;
      }
    }
    float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act3) {
      const float4 b4 = *reinterpret_cast<const float4*>(b3 + 4 * sub);
      float2 oa = lo2(b4), ob = hi2(b4);
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float4 wv = *reinterpret_cast<const float4*>(W3 + u * c_out + 4 * sub);
        oa = fma2(dup2(hf[u]), lo2(wv), oa);
        ob = fma2(dup2(hf[u]), hi2(wv), ob);
      }
      o4 = make_float4(oa.x, oa.y, ob.x, ob.y);
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

// ---- tcgen05 / tensor-memory helpers (used by the per-edge MLP of the per-graph kernel and by final_edge_tc_kernel below)
DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DEVINL float tf32_rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

DEVINL uint64_t umma_bdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version of sm_100; layout type 0 = no swizzle
  return d;
}

DEVINL void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

DEVINL void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr), "r"(__float_as_uint(a)),
               "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d))
               : "memory");
}

DEVINL void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}

DEVINL void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {      // results are valid after tmem_ld_wait()
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

DEVINL void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// bounded wait (a lost arrival must not hang the GPU): returns false after ~1 s
DEVINL bool mbar_wait(uint32_t mbar, uint32_t parity) {
  for (int it = 0; it < (1 << 24); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(mbar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}

DEVINL void tc_sync_before_mma() {         // tcgen05.st / ld of every thread ordered before the next MMA
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}


// ---------------------------------------------------------------------------------------
// Warp-level tensor-core pieces: mma.sync m16n8k8 TF32 with the 3xTF32 split
//   a*b ~= lo(a) hi(b) + hi(a) lo(b) + hi(a) hi(b)      (hi = round-to-nearest TF32, lo = TF32 of the remainder)
// fp32 accumulation; measured |err| 2.7e-7 on unit-range operands (tools/mma_probe.cu), the same as an fp32 FMA chain.
// Fragment layout (g = lane >> 2, t = lane & 3):
//   A 16x8: a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)      B 8x8: b0 (k = t, n = g)  b1 (k = t+4, n = g)
//   C 16x8: c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)
// A C tile feeds the next layer's A operand without any data movement when the next layer's k index is *defined* as
// k-step s: column t <-> unit 8s + 2t, column t+4 <-> unit 8s + 2t + 1 (the sum over k does not care about the order);
// the B fragments of that layer are gathered with the same permutation.
// ---------------------------------------------------------------------------------------
struct Split { uint32_t hi, lo; };

// cvt.rna.tf32.f32 is emulated on sm_100a (add, NaN/Inf test, select, mask = 4 instructions per conversion); the tensor
// core ignores the 13 low mantissa bits of a TF32 operand, so the split needs no explicit conversion at all:
//   mode 2 (default): hi = v with the low 13 bits masked (truncation; v - hi is then exact), lo = v - hi handed to the
//           tensor core as it is (the hardware truncates it): 2 instructions per value.  The truncation of lo is one-sided
//           (|operand| is under-estimated by < 2^-20 relative): measured error of the raw scores against the reference
//           fixtures 1e-6 .. 4.4e-6 rel-L2 (bar: 1e-4), thresholded 1000-step outputs unchanged (tools/err_probe.py)
//   mode 3 (the separate library libgdss_b200_tf32.so, GDSS_PRECISION=tf32): single-pass TF32, one MMA per product, no lo
//           term -- the reduced-precision variant; raw scores within 3e-2 rel-L2 of the reference (tested), not the default
//   mode 1: lo rounded to nearest (+ half an ulp before the hardware truncation): 3 instructions
//   mode 0: hi and lo both rounded to nearest: 4 instructions, 0.7e-6 .. 1.6e-6; 4 % slower end to end than mode 2
// (Inf stays Inf, NaN stays NaN in every mode.)
#ifndef GDSS_SPLIT_MODE
#define GDSS_SPLIT_MODE 2
#endif
DEVINL Split split_tf32(float v) {
  Split s;
#if GDSS_SPLIT_MODE == 0
  s.hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
  s.lo = __float_as_uint(v - __uint_as_float(s.hi)) + 0x1000u;
#elif GDSS_SPLIT_MODE == 1
  s.hi = __float_as_uint(v) & 0xffffe000u;                       // truncated hi: v - hi is exact
  s.lo = __float_as_uint(v - __uint_as_float(s.hi)) + 0x1000u;   // lo rounded to nearest by the add + hardware truncation
#elif GDSS_SPLIT_MODE == 2
  s.hi = __float_as_uint(v) & 0xffffe000u;
  s.lo = __float_as_uint(v - __uint_as_float(s.hi));             // lo truncated by the tensor core
#else
  s.hi = __float_as_uint(v) + 0x1000u;                           // single-pass TF32: round to nearest (add + hardware truncation)
  s.lo = 0u;
#endif
  return s;
}

DEVINL void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One output tile: c += A B with both operands split, smallest terms first.  (Three independent accumulator chains per
// tile were measured 1 % slower than this dependent chain: the kernel is bound by issue slots, not by mma latency.)
struct Acc { float m[4]; };

DEVINL void acc_init(Acc& a, float b0, float b1) {
  a.m[0] = a.m[2] = b0;
  a.m[1] = a.m[3] = b1;
}

DEVINL void mma3(Acc& c, const uint32_t (&ah)[4], const uint32_t (&al)[4], const Split& b0, const Split& b1) {
#if GDSS_SPLIT_MODE != 3
  mma_tf32(c.m, al, b0.hi, b1.hi);
  mma_tf32(c.m, ah, b0.lo, b1.lo);
#endif
  mma_tf32(c.m, ah, b0.hi, b1.hi);
}

DEVINL float acc_get(const Acc& c, int q) { return c.m[q]; }

// per-pixel MLP of one AttentionLayer (attention.py:109-114) on the tensor cores: one warp = 16 pixels per tile, the
// three layers chained in registers (C fragment -> next A fragment), weights held as split B fragments.
//   in = [M_0..M_{C-1}, P_0..P_{C-1}] -> 16 -> (16) -> c_out ; out = 2 * mlp * flag_i * flag_j
// pout may alias pin: a warp reads all inputs of its 16 pixels before it writes them (mma.sync is warp-synchronous).
template <int NLIN>
DEVINL void edge_mlp_mma(const float* M, int MS, const float* pin, int PM, const unsigned* pix, const float* fl, const float* wm,
                         const FLayer& L, int P, float* pout, float* gout, int64_t gstride, int warp, int nwarps, int lane) {
  // kmask: bit k set = input k of the first linear is live (maps of dead Q/K chains are identically zero: not gathered,
  // and a k-step made of dead inputs only is not issued)
  const unsigned kmask = (L.qk_live & ((1u << L.c_in) - 1u)) | (((1u << L.c_in) - 1u) << L.c_in);
  // (handing the tiles out through a shared counter, so that the node-MLP warps could join once their rows are done, was
  // measured 3 % slower: per-warp weight-fragment set-up and an atomic + shuffle per tile cost more than the imbalance)
  const int g = lane >> 2, t = lane & 3;
  const int C = L.c_in, c_out = L.c_out, K1 = 2 * C;
  const float* W1 = wm + L.ew_o[0];                 // [2C][16]
  const float* W2 = wm + L.ew_o[1];                 // [16][16] (NLIN == 3)
  const float* W3 = wm + L.ew_o[NLIN - 1];          // [16][c_out]
  const float* b1 = wm + L.eb_o[0];
  const float* b2 = wm + L.eb_o[1];
  const float* b3 = wm + L.eb_o[NLIN - 1];
  // ---- B fragments (split once per call)
  Split w1[2][2][2], w2[2][2][2], w3[2][2];          // [k-step][n-tile][b0 | b1]
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k0 = 8 * s + t, k1 = 8 * s + t + 4;             // layer 1: natural k order (inputs come from memory)
      w1[s][j][0] = split_tf32(k0 < K1 ? W1[k0 * 16 + 8 * j + g] : 0.f);
      w1[s][j][1] = split_tf32(k1 < K1 ? W1[k1 * 16 + 8 * j + g] : 0.f);
      if (NLIN == 3) {
        w2[s][j][0] = split_tf32(W2[(8 * s + 2 * t) * 16 + 8 * j + g]);          // permuted k order (see above)
        w2[s][j][1] = split_tf32(W2[(8 * s + 2 * t + 1) * 16 + 8 * j + g]);
      }
    }
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    w3[s][0] = split_tf32(g < c_out ? W3[(8 * s + 2 * t) * c_out + g] : 0.f);
    w3[s][1] = split_tf32(g < c_out ? W3[(8 * s + 2 * t + 1) * c_out + g] : 0.f);
  }
  float bia1[2][2], bia2[2][2], bia3[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    bia1[j][0] = b1[8 * j + 2 * t]; bia1[j][1] = b1[8 * j + 2 * t + 1];
    bia2[j][0] = NLIN == 3 ? b2[8 * j + 2 * t] : 0.f; bia2[j][1] = NLIN == 3 ? b2[8 * j + 2 * t + 1] : 0.f;
  }
  bia3[0] = 2 * t < c_out ? b3[2 * t] : 0.f;
  bia3[1] = 2 * t + 1 < c_out ? b3[2 * t + 1] : 0.f;
  const int ks1 = (K1 + 7) >> 3;                     // k-steps of layer 1 (1 or 2)
  bool klive[2][2], slive[2];                        // this lane's layer-1 inputs (k = 8 s + t + 4 q) / whole k-steps that are live
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    slive[s] = s < ks1 && ((kmask >> (8 * s)) & 0xffu);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int k = 8 * s + t + 4 * q;
      klive[s][q] = k < K1 && ((kmask >> k) & 1u);
    }
  }
  const int ntile = (P + 15) >> 4;
  for (int tile = warp; tile < ntile; tile += nwarps) {
    const int pa = tile * 16 + g, pb = pa + 8;
    const int pac = pa < P ? pa : P - 1, pbc = pb < P ? pb : P - 1;
    // ---- layer 1: A fragments gathered from the channel planes
    Acc c[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) acc_init(c[j], bia1[j][0], bia1[j][1]);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (slive[s]) {
        uint32_t ah[4], al[4];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int k = 8 * s + t + 4 * q;
          float va = 0.f, vb = 0.f;
          if (klive[s][q]) {
            const float* pl = k < C ? M + k * MS : pin + (k - C) * PM;
            va = pl[pac];
            vb = pl[pbc];
          }
          const Split sa = split_tf32(va), sb = split_tf32(vb);
          ah[2 * q] = sa.hi; al[2 * q] = sa.lo; ah[2 * q + 1] = sb.hi; al[2 * q + 1] = sb.lo;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) mma3(c[j], ah, al, w1[s][j][0], w1[s][j][1]);
      }
    }
    __syncwarp();                    // every lane has read its inputs (in-place update below)
    if (NLIN == 3) {
      Acc d[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) acc_init(d[j], bia2[j][0], bia2[j][1]);
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        uint32_t ah[4], al[4];
        const Split e0 = split_tf32(elu_f(acc_get(c[s], 0))), e1 = split_tf32(elu_f(acc_get(c[s], 1)));
        const Split e2 = split_tf32(elu_f(acc_get(c[s], 2))), e3 = split_tf32(elu_f(acc_get(c[s], 3)));
        ah[0] = e0.hi; al[0] = e0.lo; ah[1] = e2.hi; al[1] = e2.lo; ah[2] = e1.hi; al[2] = e1.lo; ah[3] = e3.hi; al[3] = e3.lo;
#pragma unroll
        for (int j = 0; j < 2; ++j) mma3(d[j], ah, al, w2[s][j][0], w2[s][j][1]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) c[j] = d[j];
    }
    Acc oa;
    acc_init(oa, bia3[0], bia3[1]);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      uint32_t ah[4], al[4];
      const Split e0 = split_tf32(elu_f(acc_get(c[s], 0))), e1 = split_tf32(elu_f(acc_get(c[s], 1)));
      const Split e2 = split_tf32(elu_f(acc_get(c[s], 2))), e3 = split_tf32(elu_f(acc_get(c[s], 3)));
      ah[0] = e0.hi; al[0] = e0.lo; ah[1] = e2.hi; al[1] = e2.lo; ah[2] = e1.hi; al[2] = e1.lo; ah[3] = e3.hi; al[3] = e3.lo;
      mma3(oa, ah, al, w3[s][0], w3[s][1]);
    }
    const float o[4] = {acc_get(oa, 0), acc_get(oa, 1), acc_get(oa, 2), acc_get(oa, 3)};
    // ---- (S + S^T), mask_adjs, stores: channels 2t, 2t+1 of pixels pa (o[0], o[1]) and pb (o[2], o[3])
    if (2 * t < c_out) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int p = r ? pb : pa;
        if (p < P) {
          const unsigned ij = pix[p];
          const float fm = 2.f * fl[ij >> 8] * fl[ij & 255];
          const float v0 = o[2 * r] * fm, v1 = o[2 * r + 1] * fm;
          if (pout) { pout[(2 * t) * PM + p] = v0; pout[(2 * t + 1) * PM + p] = v1; }
          if (gout) { gout[(int64_t)(2 * t) * gstride + p] = v0; gout[(int64_t)(2 * t + 1) * gstride + p] = v1; }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// The same per-pixel MLP on the 5th-generation tensor cores (tcgen05), for CTAs of at least 256 threads with hidden width 16:
// warps 0 .. 3 (128 threads), thread = pixel = tensor-memory lane, 128-pixel tiles, up to EDGE_SLOTS tiles in flight.
//   tensor memory, 64 columns per tile slot:  [0, 16) A operand hi   [16, 32) A operand lo   [32, 48) accumulator D (N = 16)
//   A operand of linear 1 = the pixel's 2C inputs (written by tcgen05.st straight from the channel planes in shared memory),
//   of linear 2 / 3 = ELU(D + b) split again and stored over the consumed operand; 3xTF32: hi.hi + lo.hi + hi.lo.
//   B operands (W^T, hi and lo copies, K-major without swizzle: [k / 4][16][k % 4]) are built per layer from the staged
//   weights into `bop` (the Q|K rows' region, dead between the attention maps and the next layer's chain): 6 x 1 KB.
// What it buys over the warp-level version: the three dependent linears of a 16-pixel tile were a ~3000-cycle chain per warp
// (HMMA -> C fragment -> ELU -> split -> A fragment), 16 tiles per graph of 23 nodes on 6 warps; here a graph's <= 4 tiles
// advance together through three MMA round trips, and there are no fragment gathers / per-warp weight fragments at all.
// All activations live in tensor memory between the linears: nothing is held in registers across the waits.
// pout may alias pin: a thread reads all inputs of its pixel before it writes them.
// Returns the updated mbarrier phase bits (bit s = parity of slot s's barrier).
// ---------------------------------------------------------------------------------------
#define EDGE_SLOTS 4
#define EDGE_BAR 12          // named barrier of the 128 per-edge threads

DEVINL void edge_tc_sync() {               // tcgen05.st / ld of the 128 threads ordered before the next MMAs
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  asm volatile("bar.sync %0, 128;\n" ::"n"(EDGE_BAR) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}

template <int NLIN>
__device__ __noinline__ uint32_t edge_mlp_tc(const float* M, int MS, const float* pin, int PM, const unsigned* pix, const float* fl,
                                             const float* wm, const FLayer& L, int P, float* pout, float* gout, int64_t gstride,
                                             int tid, float* bop, uint32_t tb, uint64_t* mbars, uint32_t phase, int maxslots) {
  const unsigned kmask = (L.qk_live & ((1u << L.c_in) - 1u)) | (((1u << L.c_in) - 1u) << L.c_in);
  const int C = L.c_in, c_out = L.c_out, K1 = 2 * C, K1P = (K1 + 7) & ~7;
  const float* W1 = wm + L.ew_o[0];                 // [2C][16]
  const float* W2 = wm + L.ew_o[1];                 // [16][16] (NLIN == 3)
  const float* W3 = wm + L.ew_o[NLIN - 1];          // [16][c_out]
  const float* b1 = wm + L.eb_o[0];
  const float* b2 = wm + L.eb_o[1];
  const float* b3 = wm + L.eb_o[NLIN - 1];
  // ---- B operands: [linear][hi | lo][256]
  for (int idx = tid; idx < 256; idx += 128) {
    const int k = idx >> 4, nn = idx & 15;
    const int o = (k >> 2) * 64 + nn * 4 + (k & 3);
    const Split s1 = split_tf32(k < K1 ? W1[k * 16 + nn] : 0.f);
    bop[o] = __uint_as_float(s1.hi);
    bop[256 + o] = __uint_as_float(s1.lo);
    if (NLIN == 3) {
      const Split s2 = split_tf32(W2[k * 16 + nn]);
      bop[512 + o] = __uint_as_float(s2.hi);
      bop[768 + o] = __uint_as_float(s2.lo);
    }
    const Split s3 = split_tf32(nn < c_out ? W3[k * c_out + nn] : 0.f);
    bop[1024 + o] = __uint_as_float(s3.hi);
    bop[1280 + o] = __uint_as_float(s3.lo);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // generic-proxy writes of the B operands -> async proxy (UMMA)
  const uint32_t tl = tb + ((uint32_t)(tid & ~31) << 16);            // this warp's 32 lanes
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint32_t LBO = 16 * 16, SBO = 128;
  const uint32_t bbase = smem_u32(bop);
  // one contraction (3xTF32) of every slot of the round: A = slot columns [0, 16) / [16, 32), D = [32, 48); ksteps of 8
  auto issue = [&](int nslot, int lin, int ksteps, unsigned live) {
    for (int s = 0; s < nslot; ++s) {
      const uint32_t base = tb + 64 * s;
      uint32_t acc = 0;
#pragma unroll
      for (int ps = 0; ps < (GDSS_SPLIT_MODE == 3 ? 1 : 3); ++ps) {
        const uint32_t acol = base + (ps == 1 ? 16 : 0);
        const uint32_t wb = bbase + (uint32_t)(lin * 512 + (ps == 2 ? 256 : 0)) * 4u;
        for (int ks = 0; ks < ksteps; ++ks) {
          if (!((live >> (8 * ks)) & 0xffu)) continue;               // a k-step of dead inputs only
          umma_tf32_ts(base + 32, acol + ks * 8, umma_bdesc(wb + ks * 2 * LBO, LBO, SBO), idesc, acc);
          acc = 1;
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(mbars + s))
                   : "memory");
    }
  };
  // D + bias -> ELU -> hi / lo -> the slot's A columns
  auto elu_stage = [&](int s, const float* bias) {
    const uint32_t base = tl + 64 * s;
    uint32_t dr[2][8];
    tmem_ld8_nowait(base + 32, dr[0]);
    tmem_ld8_nowait(base + 40, dr[1]);
    tmem_ld_wait();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float h[8], l[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const Split sp = split_tf32(elu_f(__uint_as_float(dr[u][q]) + bias[8 * u + q]));
        h[q] = __uint_as_float(sp.hi);
        l[q] = __uint_as_float(sp.lo);
      }
      tmem_st8(base + 8 * u, h);
      if (GDSS_SPLIT_MODE != 3) tmem_st8(base + 16 + 8 * u, l);
    }
  };
  auto wait_slot = [&](int s) {
    mbar_wait(smem_u32(mbars + s), (phase >> s) & 1u);     // bounded: a lost arrival shows up as a wrong result, not as a hang
    phase ^= 1u << s;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  };
  const int ntile = (P + 127) >> 7;
  for (int t0 = 0; t0 < ntile; t0 += maxslots) {
    const int nslot = ntile - t0 < maxslots ? ntile - t0 : maxslots;
    // ---- linear 1: the pixel's inputs -> A operand
    for (int s = 0; s < nslot; ++s) {
      const int p = (t0 + s) * 128 + tid;
      const bool valid = p < P;
      const int pc = valid ? p : P - 1;
      const uint32_t base = tl + 64 * s;
      for (int k0 = 0; k0 < K1P; k0 += 4) {
        float h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = k0 + q;
          float v = 0.f;
          if (valid && k < K1 && ((kmask >> k) & 1u)) v = k < C ? M[k * MS + pc] : pin[(k - C) * PM + pc];
          const Split sp = split_tf32(v);
          h[q] = __uint_as_float(sp.hi);
          l[q] = __uint_as_float(sp.lo);
        }
        tmem_st4(base + k0, h[0], h[1], h[2], h[3]);
        if (GDSS_SPLIT_MODE != 3) tmem_st4(base + 16 + k0, l[0], l[1], l[2], l[3]);
      }
    }
    edge_tc_sync();                        // (first round: also orders the B operands' writes before the MMAs)
    if (tid == 0) issue(nslot, 0, K1P >> 3, kmask);
    __syncwarp();                          // warp 0 reconverges before its .sync.aligned tensor-memory loads
    for (int s = 0; s < nslot; ++s) {
      wait_slot(s);
      elu_stage(s, b1);
    }
    if (NLIN == 3) {
      edge_tc_sync();
      if (tid == 0) issue(nslot, 1, 2, 0xffffu);
      __syncwarp();                          // warp 0 reconverges before its .sync.aligned tensor-memory loads
      for (int s = 0; s < nslot; ++s) {
        wait_slot(s);
        elu_stage(s, b2);
      }
    }
    edge_tc_sync();
    if (tid == 0) issue(nslot, 2, 2, 0xffffu);
    __syncwarp();                          // warp 0 reconverges before its .sync.aligned tensor-memory loads
    // ---- (S + S^T), mask_adjs, stores
    for (int s = 0; s < nslot; ++s) {
      wait_slot(s);
      float d[8];
      tmem_ld8(tl + 64 * s + 32, d);
      const int p = (t0 + s) * 128 + tid;
      if (p < P) {
        const unsigned ij = pix[p];
        const float fm = 2.f * fl[ij >> 8] * fl[ij & 255];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          if (o < c_out) {
            const float v = (d[o] + b3[o]) * fm;
            if (pout) pout[o * PM + p] = v;
            if (gout) gout[(int64_t)o * gstride + p] = v;
          }
        }
      }
    }
    // the next round's tcgen05.st into the accumulator-free columns may start at once; its MMAs are issued after the next
    // edge_tc_sync(), which also orders this round's tcgen05.ld before them
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  return phase;
}

// node MLP `multi_channel` of one AttentionLayer (attention.py:90, :106-107) on the tensor cores: one warp = 16 rows,
//   cat_c V_c[i] (K = c_in * nhid) -> 16 (ELU) -> nhid, mask, tanh (twice in X_GMH, ScoreNetwork_X.py:81)
// both layers chained in registers.  V rows have stride ldv = K + 4 (bank-conflict-free A fragments).
DEVINL void node_mlp_mma(const float* V, int ldv, int K, const float* W1, const float* b1, const float* W2, const float* b2, int n,
                         int h, const float* fl, bool twice, float* xo, int ldo, int warp, int nwarps, int lane,
                         unsigned long long kstep_live) {
  // kstep_live: bit s set = k-step s (inputs 8 s .. 8 s + 7 of the concatenated V rows) belongs to a live channel
  const int g = lane >> 2, t = lane & 3;
  const int mt = (n + 15) >> 4;
  for (int m = warp; m < mt; m += nwarps) {
    const int ra = 16 * m + g, rb = ra + 8;
    const float* va = V + (ra < n ? ra : n - 1) * ldv;
    const float* vb = V + (rb < n ? rb : n - 1) * ldv;
    Acc c[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) acc_init(c[j], b1[8 * j + 2 * t], b1[8 * j + 2 * t + 1]);
#pragma unroll 2
    for (int k0 = 0; k0 < K; k0 += 8) {
      if (!((kstep_live >> (k0 >> 3)) & 1ull)) continue;
      uint32_t ah[4], al[4];
      const Split a0 = split_tf32(va[k0 + t]), a1 = split_tf32(vb[k0 + t]);
      const Split a2 = split_tf32(va[k0 + t + 4]), a3 = split_tf32(vb[k0 + t + 4]);
      ah[0] = a0.hi; al[0] = a0.lo; ah[1] = a1.hi; al[1] = a1.lo; ah[2] = a2.hi; al[2] = a2.lo; ah[3] = a3.hi; al[3] = a3.lo;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const Split w0 = split_tf32(W1[(k0 + t) * 16 + 8 * j + g]), w1 = split_tf32(W1[(k0 + t + 4) * 16 + 8 * j + g]);
        mma3(c[j], ah, al, w0, w1);
      }
    }
    uint32_t eh[2][4], el[2][4];                       // ELU(h1) as the A fragments of the second linear (permuted k order)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const Split e0 = split_tf32(elu_f(acc_get(c[s], 0))), e1 = split_tf32(elu_f(acc_get(c[s], 1)));
      const Split e2 = split_tf32(elu_f(acc_get(c[s], 2))), e3 = split_tf32(elu_f(acc_get(c[s], 3)));
      eh[s][0] = e0.hi; el[s][0] = e0.lo; eh[s][1] = e2.hi; el[s][1] = e2.lo;
      eh[s][2] = e1.hi; el[s][2] = e1.lo; eh[s][3] = e3.hi; el[s][3] = e3.lo;
    }
    const float fa = ra < n ? fl[ra] : 0.f, fb = rb < n ? fl[rb] : 0.f;
    for (int o0 = 0; o0 < h; o0 += 8) {
      Acc d;
      acc_init(d, b2[o0 + 2 * t], b2[o0 + 2 * t + 1]);
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const Split w0 = split_tf32(W2[(8 * s + 2 * t) * h + o0 + g]), w1 = split_tf32(W2[(8 * s + 2 * t + 1) * h + o0 + g]);
        mma3(d, eh[s], el[s], w0, w1);
      }
      float v[4] = {acc_get(d, 0) * fa, acc_get(d, 1) * fa, acc_get(d, 2) * fb, acc_get(d, 3) * fb};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[q] = tanh_fast(v[q]);
        if (twice) v[q] = tanh_fast(v[q]);
      }
      if (ra < n) *reinterpret_cast<float2*>(xo + ra * ldo + o0 + 2 * t) = make_float2(v[0], v[1]);
      if (rb < n) *reinterpret_cast<float2*>(xo + rb * ldo + o0 + 2 * t) = make_float2(v[2], v[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// channel_pipeline: ONE WARP runs the whole per-channel chain of an AttentionLayer for channel c
//   degrees -> A^ x -> Q|K|V -> attention map
// with only __syncwarp() between the steps (the channels of a layer are independent until the per-edge MLP,
// attention.py:97-105), so the warps of a CTA never wait for each other inside the chain.
// ---------------------------------------------------------------------------------------
template <bool WS, int HD>
DEVINL void channel_pipeline(int lane, const float* pl, float* dv, const float* x, int ldx, int n, int f, float* axc,
                             const float* Wc, int qkvw, const float* bc, float* qkc, int QKS, float* vcol, int ldv, int h,
                             bool need_edge, bool need_node, const unsigned* pix, int P, float scale, float* Mc) {
  constexpr int AT = 4 * HD;
  // ---- degrees (layers.py:72-78): row i walked as (j, i), j < i, then the contiguous run (i, j), j > i
  for (int i = lane; i < n; i += 32) {
    float s = 1.f;
    int idx = i - 1;
    for (int j = 0; j < i; ++j) {
      s += pl[idx];
      idx += n - j - 2;
    }
    idx = i * (2 * n - i - 1) / 2;
    for (int j = i + 1; j < n; ++j) s += pl[idx++];
    dv[i] = 1.0f / sqrtf(fmaxf(s, 1.0f));
  }
  __syncwarp();
  // ---- A^ x : items (row i, 8-column group g)
  const int f8 = (f + 7) >> 3, ldax = 8 * f8;
  {
    const FastDiv df(f8);
    for (int item = lane; item < n * f8; item += 32) {
      const int i = df.div(item), g = item - i * f8;
      const float* xg = x + 8 * g;
      const float di = dv[i];
      float2 acc[4];
      {
        const float4 a = *reinterpret_cast<const float4*>(xg + i * ldx);
        const float4 b = *reinterpret_cast<const float4*>(xg + i * ldx + 4);
        const float2 d2 = dup2(di);
        acc[0] = __fmul2_rn(d2, lo2(a)); acc[1] = __fmul2_rn(d2, hi2(a));
        acc[2] = __fmul2_rn(d2, lo2(b)); acc[3] = __fmul2_rn(d2, hi2(b));
      }
      int idx = i - 1;
      for (int j = 0; j < i; ++j) {
        const float2 w = dup2(pl[idx] * dv[j]);
        idx += n - j - 2;
        const float4 a = *reinterpret_cast<const float4*>(xg + j * ldx);
        const float4 b = *reinterpret_cast<const float4*>(xg + j * ldx + 4);
        acc[0] = fma2(w, lo2(a), acc[0]); acc[1] = fma2(w, hi2(a), acc[1]);
        acc[2] = fma2(w, lo2(b), acc[2]); acc[3] = fma2(w, hi2(b), acc[3]);
      }
      idx = i * (2 * n - i - 1) / 2;
      for (int j = i + 1; j < n; ++j) {
        const float2 w = dup2(pl[idx++] * dv[j]);
        const float4 a = *reinterpret_cast<const float4*>(xg + j * ldx);
        const float4 b = *reinterpret_cast<const float4*>(xg + j * ldx + 4);
        acc[0] = fma2(w, lo2(a), acc[0]); acc[1] = fma2(w, hi2(a), acc[1]);
        acc[2] = fma2(w, lo2(b), acc[2]); acc[3] = fma2(w, hi2(b), acc[3]);
      }
      float* o = axc + i * ldax + 8 * g;
      *reinterpret_cast<float4*>(o) = make_float4(acc[0].x * di, acc[0].y * di, acc[1].x * di, acc[1].y * di);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[2].x * di, acc[2].y * di, acc[3].x * di, acc[3].y * di);
    }
  }
  __syncwarp();
  // ---- Q | K | V = (A^ x) W + b : items (row pair, 4-column group); V goes to column block c of the [i][c][h] rows
  {
    const int o_lo = need_edge ? 0 : 2 * AT, o_hi = need_node ? qkvw : 2 * AT;
    const int ogn = (o_hi - o_lo) >> 2, half = (n + 1) >> 1;
    const int K4 = (f + 3) & ~3;
    const FastDiv dog(ogn);
    for (int item = lane; item < half * ogn; item += 32) {
      const int r0 = dog.div(item), og = item - r0 * ogn;
      const int r1 = r0 + half, r1c = r1 < n ? r1 : r0;
      const int o = o_lo + 4 * og;
      const float4 bv = ldw4<WS>(bc + o);
      float2 a00 = lo2(bv), a01 = hi2(bv), a10 = a00, a11 = a01;
      const float* in0 = axc + r0 * ldax;
      const float* in1 = axc + r1c * ldax;
      const float* wp = Wc + o;
      for (int k0 = 0; k0 < K4; k0 += 4) {
        const float4 w0 = ldw4<WS>(wp + (k0 + 0) * qkvw), w1 = ldw4<WS>(wp + (k0 + 1) * qkvw);
        const float4 w2 = ldw4<WS>(wp + (k0 + 2) * qkvw), w3 = ldw4<WS>(wp + (k0 + 3) * qkvw);
        const float4 x0 = *reinterpret_cast<const float4*>(in0 + k0);
        const float4 x1 = *reinterpret_cast<const float4*>(in1 + k0);
        a00 = fma2(dup2(x0.x), lo2(w0), a00); a01 = fma2(dup2(x0.x), hi2(w0), a01);
        a10 = fma2(dup2(x1.x), lo2(w0), a10); a11 = fma2(dup2(x1.x), hi2(w0), a11);
        a00 = fma2(dup2(x0.y), lo2(w1), a00); a01 = fma2(dup2(x0.y), hi2(w1), a01);
        a10 = fma2(dup2(x1.y), lo2(w1), a10); a11 = fma2(dup2(x1.y), hi2(w1), a11);
        a00 = fma2(dup2(x0.z), lo2(w2), a00); a01 = fma2(dup2(x0.z), hi2(w2), a01);
        a10 = fma2(dup2(x1.z), lo2(w2), a10); a11 = fma2(dup2(x1.z), hi2(w2), a11);
        a00 = fma2(dup2(x0.w), lo2(w3), a00); a01 = fma2(dup2(x0.w), hi2(w3), a01);
        a10 = fma2(dup2(x1.w), lo2(w3), a10); a11 = fma2(dup2(x1.w), hi2(w3), a11);
      }
      if (o < 2 * AT) {
        *reinterpret_cast<float4*>(qkc + r0 * QKS + o) = make_float4(a00.x, a00.y, a01.x, a01.y);
        if (r1 < n) *reinterpret_cast<float4*>(qkc + r1 * QKS + o) = make_float4(a10.x, a10.y, a11.x, a11.y);
      } else {
        *reinterpret_cast<float4*>(vcol + r0 * ldv + (o - 2 * AT)) = make_float4(a00.x, a00.y, a01.x, a01.y);
        if (r1 < n) *reinterpret_cast<float4*>(vcol + r1 * ldv + (o - 2 * AT)) = make_float4(a10.x, a10.y, a11.x, a11.y);
      }
    }
  }
  __syncwarp();
  // ---- attention map (attention.py:35-49): mean over the 4 heads and both orientations of tanh(Q_h.K_h / sqrt(out_dim))
  if (need_edge) {
    for (int p = lane; p < P; p += 32) {
      const unsigned ij = pix[p];
      const int i = ij >> 8, j = ij & 255;
      const float* qi = qkc + i * QKS;
      const float* qj = qkc + j * QKS;
      float m = 0.f;
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        float2 p1 = make_float2(0.f, 0.f), p2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          const float4 a1 = *reinterpret_cast<const float4*>(qi + hh * HD + d);
          const float4 b1 = *reinterpret_cast<const float4*>(qj + AT + hh * HD + d);
          const float4 a2 = *reinterpret_cast<const float4*>(qj + hh * HD + d);
          const float4 b2 = *reinterpret_cast<const float4*>(qi + AT + hh * HD + d);
          p1 = fma2(lo2(a1), lo2(b1), p1); p1 = fma2(hi2(a1), hi2(b1), p1);
          p2 = fma2(lo2(a2), lo2(b2), p2); p2 = fma2(hi2(a2), hi2(b2), p2);
        }
        m += tanh_fast((p1.x + p1.y) * scale) + tanh_fast((p2.x + p2.y) * scale);
      }
      Mc[p] = m * 0.125f;
    }
  }
}

// channel_chain_mma: the same per-channel chain with the two contractions on the tensor cores (mma.sync 3xTF32):
//   degrees -> C1 = A~ (d . x) -> [scale rows by d: A^ x, kept in registers as the next A operand] -> Q|K|V = (A^ x) W + b -> map
// A~ is gathered from the strict-upper-triangle plane through the index table tab[i][j] (diagonal -> the 1.0 slot,
// padding -> the 0.0 slot), so the planes keep the layout the per-edge MLP reads and writes.  f <= 32.
template <bool WS, int HD, int NF, int MB>       // NF: compile-time bound of the feature tiles (f <= 8 NF); MB: row tiles per pass
__device__ __noinline__ void channel_chain_mma(int lane, const float* pl, const unsigned short* tab, int TC, float* dv, const float* x, int ldx, int n,
                              int f, int f_in_p, const float* Wc, int qkvw, const float* bc, float* qkc, int QKS, float* vcol, int ldv,
                              bool need_edge, bool need_node, const unsigned* pix, int P, float scale, float* Mc, int part, int nparts,
                              int bar_id, bool do_map) {
  // part / nparts: the warps of one channel split the row-tile passes and the pixels of the map; they meet at the named
  // barrier bar_id once Q|K are complete (each computes the degrees itself: identical values)
  constexpr int AT = 4 * HD;
  const int g = lane >> 2, t = lane & 3;
  const int ks = (n + 7) >> 3, mt = (n + 15) >> 4, nf = (f + 7) >> 3;      // k-steps over nodes, row tiles, feature tiles (<= 4)
  // ---- degrees (layers.py:72-78): row sums of A~ (self loop included through the table)
  for (int i = lane; i < n; i += 32) {
    const unsigned short* tr = tab + i * TC;
    float s0 = 0.f, s1 = 0.f;
    for (int j = 0; j < 8 * ks; j += 4) {
      const uint2 tt = *reinterpret_cast<const uint2*>(tr + j);
      s0 += pl[tt.x & 0xffff]; s1 += pl[tt.x >> 16];
      s0 += pl[tt.y & 0xffff]; s1 += pl[tt.y >> 16];
    }
    dv[i] = 1.0f / sqrtf(fmaxf(s0 + s1, 1.0f));
  }
  __syncwarp();
  int o_lo = need_edge ? 0 : 2 * AT, o_hi = need_node ? qkvw : 2 * AT;
  const float scale2 = scale * 2.8853900817779268f;
  // Work split of the warps of one channel: row-tile groups first; warps beyond the row groups split the Q|K|V output
  // column tiles of a row group (each recomputes that group's A^ x: cheaper than idling at the barrier below).
  int rp = part, rstride = nparts;
  {
    const int nrow = (mt + MB - 1) / MB;
    if (nparts > nrow) {
      const int og = nparts / nrow, op = part / nrow;
      rp = part - op * nrow;
      rstride = nrow;
      if (op >= og) rp = mt;                               // spare warp: no row group
      else {
        const int T = (o_hi - o_lo) >> 3, lo = o_lo;
        o_lo = lo + 8 * ((op * T) / og);
        o_hi = lo + 8 * (((op + 1) * T) / og);
      }
    }
  }
  for (int m0 = rp * MB; m0 < mt; m0 += MB * rstride) {
    float c1[MB][NF][4];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
      for (int jn = 0; jn < NF; ++jn)
#pragma unroll
        for (int q = 0; q < 4; ++q) c1[mb][jn][q] = 0.f;
    for (int s = 0; s < ks; ++s) {
      const int j0 = 8 * s + t, j1 = j0 + 4;
      const float d0 = j0 < n ? dv[j0] : 0.f, d1 = j1 < n ? dv[j1] : 0.f;
      const float* x0r = x + (j0 < n ? j0 : 0) * ldx + g;
      const float* x1r = x + (j1 < n ? j1 : 0) * ldx + g;
      Split b0[NF], b1[NF];
#pragma unroll
      for (int jn = 0; jn < NF; ++jn) {
        b0[jn] = split_tf32(jn < nf ? x0r[8 * jn] * d0 : 0.f);
        b1[jn] = split_tf32(jn < nf ? x1r[8 * jn] * d1 : 0.f);
      }
      uint32_t ah[MB][4], al[MB][4];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        const int mrow = m0 + mb < mt ? m0 + mb : m0;            // (a missing second tile repeats the first; never stored)
        const unsigned short* ta = tab + (16 * mrow + g) * TC;
        const unsigned short* tb = ta + 8 * TC;
        const Split a0 = split_tf32(pl[ta[j0]]), a1 = split_tf32(pl[tb[j0]]), a2 = split_tf32(pl[ta[j1]]), a3 = split_tf32(pl[tb[j1]]);
        ah[mb][0] = a0.hi; al[mb][0] = a0.lo; ah[mb][1] = a1.hi; al[mb][1] = a1.lo;
        ah[mb][2] = a2.hi; al[mb][2] = a2.lo; ah[mb][3] = a3.hi; al[mb][3] = a3.lo;
      }
      // term-major issue order: consecutive mma.sync hit different accumulators
#pragma unroll
      for (int term = GDSS_SPLIT_MODE == 3 ? 2 : 0; term < 3; ++term)
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
#pragma unroll
          for (int jn = 0; jn < NF; ++jn)
            mma_tf32(c1[mb][jn], term == 0 ? al[mb] : ah[mb], term == 1 ? b0[jn].lo : b0[jn].hi, term == 1 ? b1[jn].lo : b1[jn].hi);
    }
    // rows scaled by d_i: A^ x, re-used as the A operand of the Q|K|V contraction (k-step jn: column t <-> feature
    // 8 jn + 2t, column t+4 <-> feature 8 jn + 2t + 1)
    uint32_t qh[MB][NF][4], ql[MB][NF][4];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
      const int ra = 16 * (m0 + mb) + g, rb = ra + 8;
      const float da = ra < n ? dv[ra] : 0.f, db = rb < n ? dv[rb] : 0.f;
#pragma unroll
      for (int jn = 0; jn < NF; ++jn) {
        const Split e0 = split_tf32(c1[mb][jn][0] * da), e1 = split_tf32(c1[mb][jn][1] * da);
        const Split e2 = split_tf32(c1[mb][jn][2] * db), e3 = split_tf32(c1[mb][jn][3] * db);
        qh[mb][jn][0] = e0.hi; ql[mb][jn][0] = e0.lo; qh[mb][jn][1] = e2.hi; ql[mb][jn][1] = e2.lo;
        qh[mb][jn][2] = e1.hi; ql[mb][jn][2] = e1.lo; qh[mb][jn][3] = e3.hi; ql[mb][jn][3] = e3.lo;
      }
    }
    // (fetching the weights of the next column tile one tile ahead was measured 3.5 % slower: the extra live registers spill)
    for (int o0 = o_lo; o0 < o_hi; o0 += 8) {
      const float2 bv = WS ? *reinterpret_cast<const float2*>(bc + o0 + 2 * t) : __ldg(reinterpret_cast<const float2*>(bc + o0 + 2 * t));
      Split w0[NF], w1[NF];
#pragma unroll
      for (int jn = 0; jn < NF; ++jn) {
        const int k0 = 8 * jn + 2 * t, k1 = k0 + 1;
        w0[jn] = split_tf32(k0 < f_in_p ? (WS ? Wc[k0 * qkvw + o0 + g] : __ldg(Wc + k0 * qkvw + o0 + g)) : 0.f);
        w1[jn] = split_tf32(k1 < f_in_p ? (WS ? Wc[k1 * qkvw + o0 + g] : __ldg(Wc + k1 * qkvw + o0 + g)) : 0.f);
      }
      const int oc = o0 + 2 * t;
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        if (m0 + mb < mt) {
          Acc ca;
          acc_init(ca, bv.x, bv.y);
#pragma unroll
          for (int jn = 0; jn < NF; ++jn)
            if (jn < nf) mma3(ca, qh[mb][jn], ql[mb][jn], w0[jn], w1[jn]);
          // Q columns carry the map's 2 log2(e) / sqrt(out_dim) factor (one multiply per Q value instead of one per product)
          const float qs = o0 < AT ? scale2 : 1.f;
          const float c[4] = {acc_get(ca, 0) * qs, acc_get(ca, 1) * qs, acc_get(ca, 2) * qs, acc_get(ca, 3) * qs};
          const int ra = 16 * (m0 + mb) + g, rb = ra + 8;
          if (o0 < 2 * AT) {
            if (ra < n) *reinterpret_cast<float2*>(qkc + ra * QKS + oc) = make_float2(c[0], c[1]);
            if (rb < n) *reinterpret_cast<float2*>(qkc + rb * QKS + oc) = make_float2(c[2], c[3]);
          } else {
            if (ra < n) *reinterpret_cast<float2*>(vcol + ra * ldv + (oc - 2 * AT)) = make_float2(c[0], c[1]);
            if (rb < n) *reinterpret_cast<float2*>(vcol + rb * ldv + (oc - 2 * AT)) = make_float2(c[2], c[3]);
          }
        }
      }
    }
  }
  if (!do_map) return;                 // the maps are computed CTA-wide after a barrier (maps_cta)
  if (nparts > 1) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(32 * nparts) : "memory");
  else __syncwarp();
  // ---- attention map (attention.py:35-49): mean over the 4 heads and both orientations of tanh(Q_h.K_h / sqrt(out_dim))
  if (need_edge) {
    for (int p = part * 32 + lane; p < P; p += 32 * nparts) {
      const unsigned ij = pix[p];
      const int i = ij >> 8, j = ij & 255;
      const float* qi = qkc + i * QKS;
      const float* qj = qkc + j * QKS;
      float m = 0.f;
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        float2 p1 = make_float2(0.f, 0.f), p2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          const float4 a1 = *reinterpret_cast<const float4*>(qi + hh * HD + d);
          const float4 b1 = *reinterpret_cast<const float4*>(qj + AT + hh * HD + d);
          const float4 a2 = *reinterpret_cast<const float4*>(qj + hh * HD + d);
          const float4 b2 = *reinterpret_cast<const float4*>(qi + AT + hh * HD + d);
          p1 = fma2(lo2(a1), lo2(b1), p1); p1 = fma2(hi2(a1), hi2(b1), p1);
          p2 = fma2(lo2(a2), lo2(b2), p2); p2 = fma2(hi2(a2), hi2(b2), p2);
        }
        m += sig2(p1.x + p1.y) + sig2(p2.x + p2.y);
      }
      Mc[p] = fmaf(-0.25f, m, 1.f);             // mean of the 8 tanh = 1 - (2 / 8) * sum of 1 / (exp(2 z) + 1)
    }
  }
}

// maps_cta: the attention maps of the live channels spread evenly over the warps of the CTA (items = channel x 32-pixel
// block).  Used instead of the in-chain maps when the channels of a layer do not have equal work (dead Q/K chains or V
// convolutions switched off, or a channel count that does not divide the warp count).  Q carries the 2 log2(e) / sqrt(out_dim)
// factor (channel_chain_mma).
template <int HD>
DEVINL void maps_cta(const float* QK, int QKS, int n, const unsigned* pix, int P, float* R1, int RS, const FLayer& L, int warp,
                     int nwarps, int lane) {
  constexpr int AT = 4 * HD;
  const int nblk = (P + 31) >> 5;
  const FastDiv db(nblk);
  for (int item = warp; item < L.nqk * nblk; item += nwarps) {
    const int ci = db.div(item), p = (item - ci * nblk) * 32 + lane;
    const int c = L.qkc[ci];
    if (p >= P) continue;
    const float* qkc = QK + c * n * QKS;
    const unsigned ij = pix[p];
    const int i = ij >> 8, j = ij & 255;
    const float* qi = qkc + i * QKS;
    const float* qj = qkc + j * QKS;
    float m = 0.f;
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) {
      float2 p1 = make_float2(0.f, 0.f), p2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 a1 = *reinterpret_cast<const float4*>(qi + hh * HD + d);
        const float4 b1 = *reinterpret_cast<const float4*>(qj + AT + hh * HD + d);
        const float4 a2 = *reinterpret_cast<const float4*>(qj + hh * HD + d);
        const float4 b2 = *reinterpret_cast<const float4*>(qi + AT + hh * HD + d);
        p1 = fma2(lo2(a1), lo2(b1), p1); p1 = fma2(hi2(a1), hi2(b1), p1);
        p2 = fma2(lo2(a2), lo2(b2), p2); p2 = fma2(hi2(a2), hi2(b2), p2);
      }
      m += sig2(p1.x + p1.y) + sig2(p2.x + p2.y);
    }
    R1[c * RS + p] = fmaf(-0.25f, m, 1.f);
  }
}

// second linear of `multi_channel` + mask + tanh (attention.py:106-107; tanh'd twice in X_GMH, ScoreNetwork_X.py:81):
// items (row, 4-column group) over a subset of the CTA's threads
DEVINL void node_l2(const float* h1, const float* W, const float* b, int n, int h, const float* fl, bool twice, float* xo,
                    int ldo, int tid, int nt) {
  const int og_n = h >> 2;
  const FastDiv dg(og_n);
  for (int item = tid; item < n * og_n; item += nt) {
    const int i = dg.div(item), og = item - i * og_n;
    const float4 bv = *reinterpret_cast<const float4*>(b + 4 * og);
    float2 a0 = lo2(bv), a1 = hi2(bv);
#pragma unroll
    for (int k0 = 0; k0 < 16; k0 += 4) {
      const float4 v = *reinterpret_cast<const float4*>(h1 + i * 16 + k0);
      const float4 w0 = *reinterpret_cast<const float4*>(W + (k0 + 0) * h + 4 * og);
      const float4 w1 = *reinterpret_cast<const float4*>(W + (k0 + 1) * h + 4 * og);
      const float4 w2 = *reinterpret_cast<const float4*>(W + (k0 + 2) * h + 4 * og);
      const float4 w3 = *reinterpret_cast<const float4*>(W + (k0 + 3) * h + 4 * og);
      a0 = fma2(dup2(v.x), lo2(w0), a0); a1 = fma2(dup2(v.x), hi2(w0), a1);
      a0 = fma2(dup2(v.y), lo2(w1), a0); a1 = fma2(dup2(v.y), hi2(w1), a1);
      a0 = fma2(dup2(v.z), lo2(w2), a0); a1 = fma2(dup2(v.z), hi2(w2), a1);
      a0 = fma2(dup2(v.w), lo2(w3), a0); a1 = fma2(dup2(v.w), hi2(w3), a1);
    }
    const float f = fl[i];
    float v[4] = {a0.x * f, a0.y * f, a1.x * f, a1.y * f};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[q] = tanh_fast(v[q]);
      if (twice) v[q] = tanh_fast(v[q]);
    }
    *reinterpret_cast<float4*>(xo + i * ldo + 4 * og) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

DEVINL MM mm_make(const float* in, int in_cs, int ldin, int K, const float* W, int w_cs, int ldw, const float* bias, int b_cs,
                  int C, int n, int o_lo, int o_hi, float* out0, int o0_cs, int ld0, int split, float* out1, int o1_cs,
                  int ld1, const float* rowscale, int act, int vec_out = 1) {
  MM m;
  m.in = in; m.in_cs = in_cs; m.ldin = ldin; m.K = (K + 3) & ~3; m.W = W; m.w_cs = w_cs; m.ldw = ldw; m.bias = bias; m.b_cs = b_cs;
  m.C = C; m.n = n; m.o_lo = o_lo; m.o_hi = o_hi; m.out0 = out0; m.o0_cs = o0_cs; m.ld0 = ld0; m.split = split;
  m.out1 = out1; m.o1_cs = o1_cs; m.ld1 = ld1; m.rowscale = rowscale; m.act = act; m.vec_out = vec_out;
  return m;
}

// final node MLP of the X networks (ScoreNetwork_X.py:40-44, :84-87): xs [n][fdim] -> 2f -> 2f -> F, ELU, mask_x
DEVINL void final_node_mlp(const FNet& net, const float* xs, int XSS, int n, const float* fl, float* hA, float* hB,
                           float* out_g, int F) {
  const int HP = net.Hp;
  mm_rows(mm_make(xs, 0, XSS, net.fdim, net.fw[0], 0, HP, net.fb[0], 0, 1, n, 0, HP, hA, 0, HP, HP, nullptr, 0, 0, nullptr, ACT_ELU));
  __syncthreads();
  mm_rows(mm_make(hA, 0, HP, HP, net.fw[1], 0, HP, net.fb[1], 0, 1, n, 0, HP, hB, 0, HP, HP, nullptr, 0, 0, nullptr, ACT_ELU));
  __syncthreads();
  // the F-wide result goes straight to global memory (row stride F: scalar stores)
  mm_rows(mm_make(hB, 0, HP, HP, net.fw[2], 0, net.foutp, net.fb[2], 0, 1, n, 0, net.fout, out_g, 0, F, net.fout, nullptr, 0, 0, fl,
                  ACT_NONE, 0));
  __syncthreads();
}

// ---------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------
#define PHASE(id)                                                      \
  do {                                                                 \
    if (A.prof && threadIdx.x == 0) {                                  \
      const long long t__ = clock64();                                 \
      atomicAdd(A.prof + (id), (unsigned long long)(t__ - tprev));     \
      tprev = t__;                                                     \
    }                                                                  \
  } while (0)

// TCE: the opt-in instantiation with the per-edge MLPs on tcgen05 (edge_mlp_tc; GDSS_TCEDGE=1).  It is a separate instantiation
// because the mere presence of that path cost the default one 58 % (same source otherwise, measured on the same box).
template <int NT, bool WSM, int OCC, bool TCE = false>      // WSM: layer weights staged in shared memory by TMA (else read from global / L2)
__global__ void __launch_bounds__(NT, OCC) fused_layers_kernel(const __grid_constant__ FusedArgs A) {
  extern __shared__ __align__(16) float sm[];
  __shared__ __align__(8) uint64_t mbar_q, mbar_m, mbar_e[EDGE_SLOTS];
  __shared__ uint32_t tmem_base_s;
  long long tprev = A.prof ? clock64() : 0;
  if ((int)blockIdx.x >= A.meta[2 + A.bucket]) return;
  const int b = A.order[A.meta[5 + A.bucket] + blockIdx.x];
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
  float* xs = sm + S.xs;
  float* h1 = sm + S.h1;
  float* wq = sm + S.wq;           // TMA-staged weights of the current layer: Q|K|V convolutions
  float* wm = sm + S.wm;           //                                        : per-edge MLP + multi_channel MLP
  float* P0 = sm + S.p0;
  float* PA = sm + S.pa;
  float* R1 = sm + S.r1;           // dense A (power stack), then A.x products, then attention maps
  float* QK = sm + S.qk;           // Q|K rows [c][i][2 attn + 4]   ([R1 | QK] also hosts the final node MLP's hidden rows)
  float* R2 = sm + S.r2;           // V rows [i][c][nhid]
  unsigned short* tab = reinterpret_cast<unsigned short*>(sm + S.tab);   // triangle index of (i, j); diagonal / padding -> constant slots
  const int64_t gp0 = A.poff[b];
  const int64_t xrow0 = A.xst ? A.noff[b] : 0;
  const FastDiv dn(n), dP(P);

  // per-edge MLPs on tcgen05: tensor memory for EDGE_SLOTS tiles (64 columns each), one mbarrier per slot
  const bool tc_edge = TCE && A.use_mma && nt >= 256;
  if (tc_edge && tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(64 * EDGE_SLOTS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(sm_addr(&mbar_q)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(sm_addr(&mbar_m)) : "memory");
    for (int s = 0; s < EDGE_SLOTS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(sm_addr(&mbar_e[s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  uint32_t ph_e = 0;                // parities of the slots' mbarriers (uniform over the 128 per-edge threads)
  uint32_t ph_q = 0, ph_m = 0;      // mbarrier phase parities (one completion per fetched run)
  // ---- stage the graph
  for (int i = tid; i < N; i += nt) fl[i] = A.flags[(int64_t)b * N + i];
  for (int i = tid; i < n; i += nt) {
    const int rs = i * (2 * n - i - 1) / 2;
    for (int j = i + 1; j < n; ++j) pix[rs + j - i - 1] = (unsigned)((i << 8) | j);
  }
  if (A.use_mma) {
    const int TC = S.TC, rows = 16 * ((n + 15) >> 4);
    const FastDiv dt(TC);
    for (int idx = tid; idx < rows * TC; idx += nt) {
      const int i = dt.div(idx), j = idx - i * TC;
      int v = PM - 1;                                     // padding: the 0.0 slot
      if (i < n && j < n) {
        const int lo = i < j ? i : j, hi = i < j ? j : i;
        v = i == j ? PM - 2 : lo * (2 * n - lo - 1) / 2 + hi - lo - 1;
      }
      tab[idx] = (unsigned short)v;
    }
    int cpl = 1;                                          // planes of P0 (power stack) and PA (layer outputs)
    if (A.has_a) cpl = A.a.c_init;
    if (A.has_x && A.x.type == GDSS_NET_X_GMH && A.x.c_init > cpl) cpl = A.x.c_init;
    const int npl = cpl + (S.r1 - S.pa) / PM;
    for (int c = tid; c < npl; c += nt) {
      float* plane = c < cpl ? P0 + c * PM : PA + (c - cpl) * PM;
      plane[PM - 2] = 1.f;
      plane[PM - 1] = 0.f;
    }
  }
  {
    const FastDiv dx(XS);
    for (int idx = tid; idx < n * XS; idx += nt) {
      const int i = dx.div(idx), k = idx - i * XS;
      x0[idx] = k < F ? A.x_in[((int64_t)b * N + i) * F + k] : 0.f;
    }
    const float* Ag = A.adj + (int64_t)b * N * N;
    for (int idx = tid; idx < n * n; idx += nt) {
      const int i = dn.div(idx), j = idx - i * n;
      R1[i * LD + j] = Ag[i * N + j];
    }
  }
  if (tc_edge) asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  uint32_t tmem_b = 0;
  if (tc_edge) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    tmem_b = tmem_base_s;
  }
  // ---- power stack (graph_utils.py:133-142), c_init <= 2: plane 0 = A, plane 1 = A @ A
  int c_init = 1;
  if (A.has_a) c_init = A.a.c_init;
  if (A.has_x && A.x.type == GDSS_NET_X_GMH && A.x.c_init > c_init) c_init = A.x.c_init;
  for (int p = tid; p < P; p += nt) {
    const unsigned ij = pix[p];
    const int i = ij >> 8, j = ij & 255;
    const float a = R1[i * LD + j];
    P0[p] = a;
    float a2 = 0.f;
    if (c_init > 1) {
      for (int m = 0; m < n; ++m) a2 = fmaf(R1[i * LD + m], R1[m * LD + j], a2);
      P0[PM + p] = a2;
    }
    if (A.has_a) {
      A.stackT[gp0 + p] = a;
      if (A.a.c_init > 1) A.stackT[A.pstride + gp0 + p] = a2;
    }
  }
  __syncthreads();
  PHASE(0);

  // ---- the two networks
  for (int which = 0; which < 2; ++which) {
    const bool isA = which == 0;
    if (isA ? !A.has_a : !A.has_x) continue;
    const FNet& net = isA ? A.a : A.x;
    if (!isA) {
      // node feature stack [x, x_1 .. x_depth] (ScoreNetwork_X.py:40, :84), padded columns zero
      if (A.xst) {
        for (int idx = tid; idx < n * F; idx += nt) {
          const int k = dn.div(idx), i = idx - k * n;
          A.xst[(int64_t)k * A.nstride + xrow0 + i] = x0[i * XS + k];
        }
      } else {
        const FastDiv dq(XSS);
        for (int idx = tid; idx < n * XSS; idx += nt) {
          const int i = dq.div(idx), k = idx - i * XSS;
          xs[idx] = k < F ? x0[i * XS + k] : 0.f;
        }
      }
    }
    const int h = net.nhid;
    const FastDiv dh(h);
    float* hA = R1;
    float* hB = R1 + n * net.Hp;
    if (net.type == GDSS_NET_X) {
      // ScoreNetworkX (ScoreNetwork_X.py:32-46): x_l = tanh(DenseGCNConv_l(x_{l-1}, adj)), raw adjacency only
      gcn_degrees(P0, PM, 1, n, NBP, dinv, dn);
      __syncthreads();
      const float* xin = x0;
      const int RS0 = S.RS;
      for (int l = 0; l < net.depth; ++l) {
        const int f = l == 0 ? F : h, f8 = (f + 7) >> 3;
        gcn_ax(P0, PM, dinv, NBP, xin, XS, 1, n, f8, 8 * f8, R1, RS0, dn);
        __syncthreads();
        mm_rows(mm_make(R1, 0, 8 * f8, f, net.gw[l], 0, h, net.gb[l], 0, 1, n, 0, h, xb, 0, XS, h, nullptr, 0, 0, nullptr, ACT_TANH));
        __syncthreads();
        if (A.xst) {
          for (int idx = tid; idx < n * h; idx += nt) {
            const int o = dn.div(idx), i = idx - o * n;
            A.xst[(int64_t)(F + l * h + o) * A.nstride + xrow0 + i] = xb[i * XS + o];
          }
        } else {
          for (int idx = tid; idx < n * h; idx += nt) {
            const int i = dh.div(idx), o = idx - i * h;
            xs[i * XSS + F + l * h + o] = xb[i * XS + o];
          }
        }
        xin = xb;
      }
      __syncthreads();
      if (!A.xst) final_node_mlp(net, xs, XSS, n, fl, hA, hB, A.raw_x + (int64_t)b * N * F, F);
      continue;
    }
    // AttentionLayer stack (ScoreNetwork_A.py:131-141, ScoreNetwork_X.py:75-84)
    const float* pin = P0;
    const float* xin = x0;
    int cbase = net.c_init;
    const float scale = 1.0f / sqrtf((float)h);
    __syncthreads();                 // the previous network is done with wq / wm
    if (WSM && tid == 0) {
      tma_fetch(wq, net.L[0].blk_q, net.L[0].q_floats, &mbar_q);
      tma_fetch(wm, net.L[0].blk_m, net.L[0].m_floats, &mbar_m);
    }
    const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
    const int RS = S.RS;
    for (int l = 0; l < net.depth; ++l) {
      const FLayer& L = net.L[l];
      const int C = L.c_in, f = L.f_in, f8 = (f + 7) >> 3, attn = L.attn;
      const bool more = l + 1 < net.depth;
      const int LDV = C * h + 4;       // row stride of the V rows [i][c][h] (+4: conflict-free mma A fragments)
      const float* wqp = WSM ? wq : L.blk_q;
      const bool chain_mma = A.use_mma && ((f <= 16 && attn == 16) || (f <= 32 && attn == 32));
      const int nact = L.nact;           // channels with a live Q/K chain or V convolution (all of them unless pruned)
      const bool any_qk = L.nqk > 0, any_v = L.v_live != 0;
      if (WSM) {
        mbar_wait_parity(&mbar_q, ph_q);
        ph_q ^= 1;
      }
      if (nact == 0) {
        // nothing of this layer's convolutions reaches an output
      } else if (chain_mma || 2 * C > nwarps) {
        // ---- one warp per channel, no CTA barrier inside the chain
        const int nparts = (chain_mma && nwarps >= 2 * nact) ? nwarps / nact : 1;       // warps per channel (mma chain only)
        // channels of unequal work (dead chains switched off) or a channel count that leaves warps idle: the maps are
        // computed CTA-wide after a barrier instead of inside the chains
        const bool split_maps = chain_mma && any_qk && (L.nqk != nact || (nact * nparts) % nwarps != 0);
        for (int u = warp; u < nact * nparts; u += nwarps) {
          const int ci = nparts > 1 ? u % nact : u, part = nparts > 1 ? u / nact : 0;
          const int c = L.act[ci];
          const bool ce = (L.qk_live >> c) & 1u, cn = (L.v_live >> c) & 1u;
          const float* Wc = wqp + c * L.f_in_p * L.qkvw;
          const float* bc = wqp + L.bq_o + c * L.qkvw;
          if (A.use_mma && f <= 16 && attn == 16 && (A.mb1 || nparts > 1)) {   // graphs of at most 16 nodes: one row tile;
                                                                               // several warps per channel: one tile per warp
            channel_chain_mma<WSM, 4, 2, 1>(lane, pin + c * PM, tab, S.TC, dinv + c * NBP, xin, XS, n, f, L.f_in_p, Wc, L.qkvw, bc,
                                            QK + c * n * QKS, QKS, R2 + c * h, LDV, ce, cn, pix, P, scale, R1 + c * RS, part, nparts, 2 + c, !split_maps);
          } else if (A.use_mma && f <= 16 && attn == 16) {
            channel_chain_mma<WSM, 4, 2, 2>(lane, pin + c * PM, tab, S.TC, dinv + c * NBP, xin, XS, n, f, L.f_in_p, Wc, L.qkvw, bc,
                                            QK + c * n * QKS, QKS, R2 + c * h, LDV, ce, cn, pix, P, scale, R1 + c * RS, part, nparts, 2 + c, !split_maps);
          } else if (A.use_mma && f <= 32 && attn == 32) {
            channel_chain_mma<WSM, 8, 4, 1>(lane, pin + c * PM, tab, S.TC, dinv + c * NBP, xin, XS, n, f, L.f_in_p, Wc, L.qkvw, bc,
                                         QK + c * n * QKS, QKS, R2 + c * h, LDV, ce, cn, pix, P, scale, R1 + c * RS, part, nparts, 2 + c, !split_maps);
          } else if (part > 0) {
            // (the FFMA chain runs one warp per channel)
          } else if (attn == 16)
            channel_pipeline<WSM, 4>(lane, pin + c * PM, dinv + c * NBP, xin, XS, n, f, R1 + c * RS, Wc, L.qkvw, bc, QK + c * n * QKS, QKS,
                                     R2 + c * h, LDV, h, ce, cn, pix, P, scale, R1 + c * RS);
          else
            channel_pipeline<WSM, 8>(lane, pin + c * PM, dinv + c * NBP, xin, XS, n, f, R1 + c * RS, Wc, L.qkvw, bc, QK + c * n * QKS, QKS,
                                     R2 + c * h, LDV, h, ce, cn, pix, P, scale, R1 + c * RS);
        }
        PHASE(8);                    // warp 0's own channel chain; PHASE(3) below = its wait for the slowest warp
        if (split_maps) {
          __syncthreads();           // Q | K of every live channel complete
          if (attn == 16) maps_cta<4>(QK, QKS, n, pix, P, R1, RS, L, warp, nwarps, lane);
          else maps_cta<8>(QK, QKS, n, pix, P, R1, RS, L, warp, nwarps, lane);
        }
      } else {
        // ---- few channels (first layer): every step spread over the whole CTA (all channels are evaluated when any is live)
        gcn_degrees(pin, PM, C, n, NBP, dinv, dn);
        __syncthreads();
        PHASE(1);
        gcn_ax(pin, PM, dinv, NBP, xin, XS, C, n, f8, 8 * f8, R1, RS, dn);
        __syncthreads();
        PHASE(2);
        // Q | K | V = (A^ x) W + b for every channel (attention.py:32-34); V rows are stored [i][c][h]
        mm_rows_t<WSM>(mm_make(R1, RS, 8 * f8, f, wqp, L.f_in_p * L.qkvw, L.qkvw, wqp + L.bq_o, L.qkvw, C, n,
                               any_qk ? 0 : 2 * attn, any_v ? L.qkvw : 2 * attn, QK, n * QKS, QKS, 2 * attn, R2, h, LDV,
                               nullptr, ACT_NONE));
        __syncthreads();
        if (any_qk) {
          if (attn == 16) attention_maps<4>(QK, QKS, pix, C, n, P, RS, scale, R1, dP);
          else attention_maps<8>(QK, QKS, pix, C, n, P, RS, scale, R1, dP);
        }
        PHASE(4);
      }
      // dead chains: their maps are identically zero / their V blocks are never read with a live weight -- zero both so
      // that every reader sees finite values (the regions are private to the dead channel)
      // (the mma variants never read them: edge_mlp_mma gathers live inputs only, node_mlp_mma skips dead k-steps)
      if (!A.use_mma && L.need_edge && !L.edge_const && L.nqk < C) {
        const bool full_eval = nact > 0 && !(chain_mma || 2 * C > nwarps) && any_qk;     // the CTA-wide path wrote every map
        if (!full_eval)
          for (int c = 0; c < C; ++c)
            if (!((L.qk_live >> c) & 1u))
              for (int p = tid; p < P; p += nt) R1[c * RS + p] = 0.f;
      }
      if (!A.use_mma && L.need_node && L.v_live != ((1u << C) - 1u)) {
        const bool full_eval = nact > 0 && !(chain_mma || 2 * C > nwarps) && any_v;
        if (!full_eval)
          for (int c = 0; c < C; ++c)
            if (!((L.v_live >> c) & 1u))
              for (int idx = tid; idx < n * h; idx += nt) {
                const int i = dh.div(idx), o = idx - i * h;
                R2[i * LDV + c * h + o] = 0.f;
              }
      }
      __syncthreads();
      if (WSM && more && tid == 0) tma_fetch(wq, net.L[l + 1].blk_q, net.L[l + 1].q_floats, &mbar_q);
      PHASE(3);
      // ---- node MLP (attention.py:106-107) on the last two warps, per-edge MLP (:109-114) on the others.
      // x_in is dead by now: layers >= 1 update xb in place (layer 0 reads x0, which the X network needs again)
      if (WSM) {
        mbar_wait_parity(&mbar_m, ph_m);
        ph_m ^= 1;
      }
      const float* wmp = WSM ? wm : L.blk_m;
      // split only when both exist; one node warp per 16-row tile (the mma node MLP works on whole tiles)
      // (16 warps, 33 .. 48 nodes: three row tiles -> three node warps; two would leave one of them with two tiles, the long pole)
      const int node_warps = (L.need_node && L.need_edge && nwarps >= 4)
                                 ? ((A.use_mma && (nwarps < 8 || n <= 16)) ? 1 : ((A.use_mma && nwarps >= 16 && n > 32) ? 3 : 2))
                                 : 0;
      // per-edge MLP on tcgen05 (warps 0 .. 3) when the layer has the shape the tile code is built for
      const bool edge_tc = TCE && tc_edge && L.need_edge && !L.edge_const && (L.nlin == 2 || L.nlin == 3);     // (hidden width 16, <= 8 channels: net_fusable)
      const int nt_edge = edge_tc ? 128 : nt - 32 * node_warps;
      if (L.need_node && (node_warps == 0 || tid >= nt - 32 * node_warps)) {
        const int ltid = node_warps ? tid - (nt - 32 * node_warps) : tid, lnt = node_warps ? 32 * node_warps : nt;
        if (A.use_mma) {
          node_mlp_mma(R2, LDV, C * h, wmp + L.nw0_o, wmp + L.nb0_o, wmp + L.nw1_o, wmp + L.nb1_o, n, h, fl, !isA, xb, XS, ltid >> 5,
                       lnt >> 5, lane, L.vk_live);
        } else {
          node_l1(R2, LDV, C * h, wmp + L.nw0_o, wmp + L.nb0_o, n, h1, ltid, lnt);
          if (node_warps) asm volatile("bar.sync 1, %0;" ::"r"(lnt) : "memory");
          else __syncthreads();
          node_l2(h1, wmp + L.nw1_o, wmp + L.nb1_o, n, h, fl, !isA, xb, XS, ltid, lnt);
        }
      }
      if (L.need_edge && (edge_tc ? tid < 128 : (node_warps == 0 || tid < nt_edge))) {
        if (!node_warps && L.need_node) __syncthreads();       // (never both: kept for symmetry of the barrier count)
        float* gout = isA ? A.stackT + (int64_t)cbase * A.pstride + gp0 : nullptr;
        float* pout = more ? PA : nullptr;
        if (L.edge_const) {
          // every input weight of `mlp` is dead: the channels are the constant mlp(0) (attention.py:109-114), symmetrised and masked
          for (int p = tid; p < P; p += nt_edge) {
            const unsigned ij = pix[p];
            const float fm = 2.f * fl[ij >> 8] * fl[ij & 255];
            for (int o = 0; o < L.c_out; ++o) {
              const float v = L.cst[o] * fm;
              if (pout) pout[o * PM + p] = v;
              if (gout) gout[(int64_t)o * A.pstride + p] = v;
            }
          }
        } else if (TCE && edge_tc) {
          // B operands in the Q|K rows' region: dead between the attention maps and the next layer's chain
          if constexpr (TCE) {
            if (L.nlin == 3) ph_e = edge_mlp_tc<3>(R1, RS, pin, PM, pix, fl, wmp, L, P, pout, gout, A.pstride, tid, QK, tmem_b, mbar_e, ph_e, A.tce_slots);
            else ph_e = edge_mlp_tc<2>(R1, RS, pin, PM, pix, fl, wmp, L, P, pout, gout, A.pstride, tid, QK, tmem_b, mbar_e, ph_e, A.tce_slots);
          }
        } else if (A.use_mma) {
          if (L.nlin == 3) edge_mlp_mma<3>(R1, RS, pin, PM, pix, fl, wmp, L, P, pout, gout, A.pstride, warp, nt_edge >> 5, lane);
          else edge_mlp_mma<2>(R1, RS, pin, PM, pix, fl, wmp, L, P, pout, gout, A.pstride, warp, nt_edge >> 5, lane);
        } else {
          if (L.nlin == 3) edge_mlp<3>(R1, RS, pin, PM, pix, fl, wmp, L, P, pout, gout, A.pstride, tid, nt_edge);
          else edge_mlp<2>(R1, RS, pin, PM, pix, fl, wmp, L, P, pout, gout, A.pstride, tid, nt_edge);
        }
      }
      PHASE(7);                      // warp 0's share of the per-edge MLP; PHASE(5) below = its wait (node warps, other edge warps)
      __syncthreads();
      if (WSM && more && tid == 0) tma_fetch(wm, net.L[l + 1].blk_m, net.L[l + 1].m_floats, &mbar_m);
      PHASE(5);
      if (L.need_node) {
        if (!isA) {
          if (A.xst) {
            for (int idx = tid; idx < n * h; idx += nt) {
              const int o = dn.div(idx), i = idx - o * n;
              A.xst[(int64_t)(F + l * h + o) * A.nstride + xrow0 + i] = xb[i * XS + o];
            }
          } else {
            for (int idx = tid; idx < n * h; idx += nt) {
              const int i = dh.div(idx), o = idx - i * h;
              xs[i * XSS + F + l * h + o] = xb[i * XS + o];
            }
          }
        }
        xin = xb;
      }
      pin = PA;
      cbase += L.c_out;
    }
    if (!isA && !A.xst) {
      __syncthreads();
      final_node_mlp(net, xs, XSS, n, fl, hA, hB, A.raw_x + (int64_t)b * N * F, F);
      PHASE(6);
    }
  }
  if (tc_edge) {
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_b), "n"(64 * EDGE_SLOTS) : "memory");
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
  int N, fdim, H, Hs;
  const int* neff; int B, F, foutp;      // node mode (X networks): rows = b * N + i, F outputs per row, w3 [Hs][foutp]
  int fb;                                // bits per index field of gmap (6: per-graph path, 9: general path)
  const int* nmap;                       // node mode: compact row -> b << 6 | i  (-1 on the padding of the last tile)
  int64_t dense_bs;                      // != 0 (general path): stackT is the dense channel stack [b][c][N][N] with this batch stride
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
    w3s[idx] = idx < H ? a.w3[idx * 4] : 0.f;      // [Hp][foutp = 4]: the single output column
  }
  for (int idx = tid; idx < K1 * HP; idx += 256) {
    int k = idx / HP, o = idx - k * HP;
    W1s[idx] = o < H ? a.w1[k * a.Hs + o] : 0.f;
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
    W2s[idx] = o < H ? a.w2[k * a.Hs + o] : 0.f;
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

// ---------------------------------------------------------------------------------------
// final_edge_tc_kernel: the same final MLP on the 5th-generation tensor cores (tcgen05), fp32-faithful through
// the 3xTF32 split  x*w ~= hi(x)hi(w) + lo(x)hi(w) + hi(x)lo(w)  (hi = round-to-nearest TF32, lo = TF32 of the
// remainder; accumulation in fp32 in tensor memory).  Persistent: one CTA per SM loops over 128-pixel tiles.
//
//   tensor memory (columns):  [0, NP)  A-operand hi      [NP, 2NP)  A-operand lo      [2NP, 3NP)  accumulator D
//     layer 1: A = the K1P channel values of the tile's 128 pixels (hi in columns [0,K1P), lo in [K1P, 2 K1P)),
//              written by tcgen05.st straight from registers (thread = pixel = TMEM lane);
//     layer 2: A = ELU(D + b1), split again and stored over the dead layer-1 operand.
//   shared memory: W1^T, W2^T as B operands, K-major without swizzle ([k/4][n][k%4], hi and lo copies),
//              staged once per CTA; biases and the last layer's weight vector.
//   256 threads: warp w works on TMEM lanes 32 (w % 4) .. +31 and on column half w / 4.
// ---------------------------------------------------------------------------------------
// (Measured and dropped: issuing the second contraction in two K chunks so that its first half overlaps the conversion of
// the second half of the layer-2 operand -- the extra CTA barrier per tile cost more than the overlap gained, +5 %.)
// NODE: final MLP of an X network (FOP = padded output width), rows = nodes.  NS: column splits = warps per TMEM lane
// quarter (128 NS threads).  NS = 4 did not help (the tile time is the serial sum of the MMAs and the epilogues' ALU work,
// not latency): every instance uses NS = 2.
template <int NP, int K1P, int FOP, bool NODE, int NS>
__global__ void __launch_bounds__(128 * NS, 1) final_edge_tc_kernel(FinalTriArgs a, int* __restrict__ err) {
  extern __shared__ __align__(128) float sm[];
  constexpr int NT = 128 * NS;
  constexpr int KH = K1P / NS;            // layer-1 operand columns per column split
  constexpr int CH = NP / NS;             // hidden columns per column split
  static_assert(NP % 16 == 0 && K1P % 8 == 0 && K1P <= NP && 3 * NP <= 512 && CH % 8 == 0 && KH % 4 == 0, "shape");
  float* W1hi = sm;                       // [K1P/4][NP][4]
  float* W1lo = W1hi + K1P * NP;
  float* W2hi = W1lo + K1P * NP;          // [NP/4][NP][4]
  float* W2lo = W2hi + NP * NP;
  float* b1s = W2lo + NP * NP;            // [NP]
  float* b2s = b1s + NP;
  float* w3s = b2s + NP;                  // [NP] (edge) / [NP][FOP] (node)
  float* part = w3s + NP * (NODE ? FOP : 1);   // [NS-1][128] ([NS-1][FOP][128]) partial dot products of column splits 1..
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar, mbar2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp >> 2;             // column split of this warp (0 .. NS-1)
  const int row = 32 * (warp & 3) + lane;
  const int K1 = a.fdim, H = a.H, N = a.N;
  // ---- one-time set-up: B operands (hi / lo), biases, tensor memory, mbarrier
  for (int idx = tid; idx < K1P * NP; idx += NT) {
    const int k = idx / NP, n = idx - k * NP;
    const float w = (k < K1 && n < H) ? a.w1[k * a.Hs + n] : 0.f;
    const float hi = tf32_rna(w);
    const int o = (k >> 2) * (NP * 4) + n * 4 + (k & 3);
    W1hi[o] = hi;
    W1lo[o] = tf32_rna(w - hi);
  }
  for (int idx = tid; idx < NP * NP; idx += NT) {
    const int k = idx / NP, n = idx - k * NP;
    const float w = (k < H && n < H) ? a.w2[k * a.Hs + n] : 0.f;
    const float hi = tf32_rna(w);
    const int o = (k >> 2) * (NP * 4) + n * 4 + (k & 3);
    W2hi[o] = hi;
    W2lo[o] = tf32_rna(w - hi);
  }
  for (int idx = tid; idx < NP; idx += NT) {
    b1s[idx] = idx < H ? a.b1[idx] : 0.f;
    b2s[idx] = idx < H ? a.b2[idx] : 0.f;
    if (!NODE) w3s[idx] = idx < H ? a.w3[idx * 4] : 0.f;
  }
  if (NODE)
    for (int idx = tid; idx < NP * FOP; idx += NT) {
      const int k = idx / FOP, o = idx - k * FOP;
      w3s[idx] = (k < H && o < a.foutp) ? a.w3[k * a.foutp + o] : 0.f;
    }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar2)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // generic-proxy writes of W -> async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tb = tmem_base_s;
  const uint32_t tlane = tb + ((uint32_t)(32 * (warp & 3)) << 16);
  // Tensor-memory columns.  When they fit (2 K1P + 4 NP <= 512) the layer-1 operand, the layer-2 operand and the two
  // accumulators have their own columns, and the first contraction of the NEXT tile is issued together with the second
  // contraction of the current one: it runs on the tensor pipe while the threads are busy with the second epilogue.
  constexpr bool PIPE = 2 * K1P + 4 * NP <= 512;
  constexpr uint32_t colA1 = 0;                                   // layer-1 operand: hi [0, K1P), lo [K1P, 2 K1P)
  constexpr uint32_t colAhi = PIPE ? 2 * K1P : 0, colAlo = colAhi + NP;     // layer-2 operand hi / lo
  constexpr uint32_t colD = colAlo + NP;                          // accumulator of the first contraction
  constexpr uint32_t colD2 = PIPE ? colD + NP : colD;             // accumulator of the second contraction
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint32_t LBO = NP * 16, SBO = 128;
  const uint32_t mb = smem_u32(&mbar), mb2 = smem_u32(&mbar2);
  const float b3 = a.b3[0];
  uint32_t phase = 0, phase2 = 0;
  bool ok = true;
  const int ntiles = NODE ? a.meta[9] : a.meta[1];
  // channel values of this thread's pixel for the first tile; every later tile is fetched one tile ahead (the loads
  // are in flight during the two contractions and epilogues of the current tile)
  float v[KH];
  // channel values of this thread's pixel of tile `tile`: channel-major planes over the batch's pixels (per-graph path) or,
  // on the general path, the dense channel stack [b][c][N][N] addressed through gmap
  auto fetch = [&](int tile) {
    const int64_t gpn = (int64_t)tile * 128 + row;
    bool valid = tile < ntiles;
    const float* src = a.stackT + gpn;
    int64_t cs = a.pstride;
    if (!NODE && a.dense_bs) {
      const int g = valid ? __ldg(a.gmap + gpn) : -1;
      valid = g >= 0;
      const int b = g >> (2 * a.fb), i = (g >> a.fb) & ((1 << a.fb) - 1), j = g & ((1 << a.fb) - 1);
      src = a.stackT + (valid ? (int64_t)b * a.dense_bs + (int64_t)i * N + j : 0);
      cs = (int64_t)N * N;
    }
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) {
      const int c = half * KH + kk;
      v[kk] = (c < K1 && valid) ? __ldg(src + (int64_t)c * cs) : 0.f;
    }
  };
  fetch(blockIdx.x);
  // layer-1 operand of the tile held in v[] -> hi / lo -> tensor memory; then v[] <- the tile `pre` (if any)
  auto stage_a1 = [&](int pre) {
#pragma unroll
    for (int kk = 0; kk < KH; kk += 4) {
      float h[4], l[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const Split sp = split_tf32(v[kk + q]);
        h[q] = __uint_as_float(sp.hi);
        l[q] = __uint_as_float(sp.lo);
      }
      tmem_st4(tlane + colA1 + half * KH + kk, h[0], h[1], h[2], h[3]);
      if (GDSS_SPLIT_MODE != 3) tmem_st4(tlane + colA1 + K1P + half * KH + kk, l[0], l[1], l[2], l[3]);
    }
    if (pre < ntiles) fetch(pre);
  };
  // first contraction (3xTF32: hi.hi + lo.hi + hi.lo) of the staged layer-1 operand -> colD, completion on `mb`
  auto issue_mma1 = [&]() {
    uint32_t acc = 0;
#pragma unroll
    for (int ps = 0; ps < (GDSS_SPLIT_MODE == 3 ? 1 : 3); ++ps) {
      const uint32_t acol = tb + colA1 + (ps == 1 ? K1P : 0);
      const uint32_t bbase = smem_u32(ps == 2 ? W1lo : W1hi);
#pragma unroll
      for (int s = 0; s < K1P / 8; ++s) {
        umma_tf32_ts(tb + colD, acol + s * 8, umma_bdesc(bbase + s * 2 * LBO, LBO, SBO), idesc, acc);
        acc = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mb) : "memory");
  };
  if (PIPE && (int)blockIdx.x < ntiles) {                    // prologue: the first tile's first contraction
    stage_a1(blockIdx.x + gridDim.x);
    tc_sync_before_mma();
    if (tid == 0) issue_mma1();
  }
  for (int tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x) {
    const int64_t gp = (int64_t)tile * 128 + row;
    const int next = tile + (int)gridDim.x;
    // where this thread's result goes and its flag mask: a chain of two dependent global loads, issued here so that it is
    // complete long before the second epilogue needs it (it used to stall the whole CTA at the end of every tile)
    int out_b = -1, out_i = 0, out_j = 0;
    float out_f = 0.f;
    if (half == 0) {
      if (NODE) {
        const int nm = __ldg(a.nmap + gp);
        if (nm >= 0) {
          out_b = nm >> 6;
          out_i = nm & 63;
          out_f = __ldg(a.flags + (int64_t)out_b * N + out_i);
        }
      } else {
        const int gv = __ldg(a.gmap + gp);
        if (gv >= 0) {
          out_b = gv >> (2 * a.fb); out_i = (gv >> a.fb) & ((1 << a.fb) - 1); out_j = gv & ((1 << a.fb) - 1);
          const float* fl = a.flags + (int64_t)out_b * N;
          out_f = __ldg(fl + out_i) * __ldg(fl + out_j);
        }
      }
    }
    if (!PIPE) {
      stage_a1(next);
      tc_sync_before_mma();
      if (tid == 0) issue_mma1();
      __syncwarp();
    }
    ok = mbar_wait(mb, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // ---- epilogue 1: ELU(D + b1) -> hi / lo -> layer-2 operand
    {
      constexpr int NB8 = CH / 8;                      // 8-column groups of this thread's column half
      constexpr int GB = NB8 % 3 == 0 ? 3 : (NB8 % 2 == 0 ? 2 : 1);      // groups per batch (one tcgen05.wait::ld per batch)
#pragma unroll
      for (int b0 = 0; b0 < NB8; b0 += GB) {
        uint32_t dr[GB][8];
#pragma unroll
        for (int u = 0; u < GB; ++u) tmem_ld8_nowait(tlane + colD + half * CH + 8 * (b0 + u), dr[u]);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < GB; ++u) {
          const int col = half * CH + 8 * (b0 + u);
          float h[8], l[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const Split sp = split_tf32(elu_f(__uint_as_float(dr[u][q]) + b1s[col + q]));
            h[q] = __uint_as_float(sp.hi);
            l[q] = __uint_as_float(sp.lo);
          }
          tmem_st8(tlane + colAhi + col, h);
          if (GDSS_SPLIT_MODE != 3) tmem_st8(tlane + colAlo + col, l);
        }
      }
    }
    if (PIPE && next < ntiles) stage_a1(next + (int)gridDim.x);     // the layer-1 columns are free: its contraction is complete
    tc_sync_before_mma();
    if (tid == 0) {
      uint32_t acc = 0;
#pragma unroll
      for (int ps = 0; ps < (GDSS_SPLIT_MODE == 3 ? 1 : 3); ++ps) {
        const uint32_t acol = tb + (ps == 1 ? colAlo : colAhi);
        const uint32_t bbase = smem_u32(ps == 2 ? W2lo : W2hi);
#pragma unroll
        for (int s = 0; s < NP / 8; ++s) {
          umma_tf32_ts(tb + colD2, acol + s * 8, umma_bdesc(bbase + s * 2 * LBO, LBO, SBO), idesc, acc);
          acc = 1;
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(PIPE ? mb2 : mb) : "memory");
      if (PIPE && next < ntiles) issue_mma1();          // next tile's first contraction: overlaps the epilogue below
    }
    __syncwarp();       // warp 0 reconverges before its .sync.aligned tensor-memory loads (lane 0 may still be issuing when the
                        // barrier of the second contraction completes; a partial-warp tcgen05.ld was observed to return garbage)
    if (PIPE) {
      ok = ok && mbar_wait(mb2, phase2);
      phase2 ^= 1;
    } else {
      ok = ok && mbar_wait(mb, phase);
      phase ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (NODE) {
      // ---- epilogue 2 (X networks, ScoreNetwork_X.py:43-46, :86-89): ELU(D + b2) W3 + b3, mask_x
      float dv[FOP];
#pragma unroll
      for (int o = 0; o < FOP; ++o) dv[o] = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < CH; c0 += 8) {
        const int col = half * CH + c0;
        float d[8];
        tmem_ld8(tlane + colD2 + col, d);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float e = elu_f(d[q] + b2s[col + q]);
#pragma unroll
          for (int o = 0; o < FOP; o += 4) {
            const float4 w = *reinterpret_cast<const float4*>(w3s + (col + q) * FOP + o);
            dv[o] = fmaf(e, w.x, dv[o]); dv[o + 1] = fmaf(e, w.y, dv[o + 1]);
            dv[o + 2] = fmaf(e, w.z, dv[o + 2]); dv[o + 3] = fmaf(e, w.w, dv[o + 3]);
          }
        }
      }
      if (half > 0) {
#pragma unroll
        for (int o = 0; o < FOP; ++o) part[((half - 1) * FOP + o) * 128 + row] = dv[o];
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      ok = __syncthreads_and(ok);             // the tile's only barrier after the MMAs: partial sums visible, uniform loop exit
      if (half == 0 && out_b >= 0) {
        {
          const float f = out_f;
          float* O = a.out + ((int64_t)out_b * N + out_i) * a.F;
#pragma unroll
          for (int o = 0; o < FOP; ++o) {
            float acc = dv[o];
#pragma unroll
            for (int q = 0; q < NS - 1; ++q) acc += part[(q * FOP + o) * 128 + row];
            if (o < a.F) O[o] = (acc + a.b3[o]) * f;
          }
        }
      }
      continue;                               // (part[] is rewritten only after the next tile's barrier before its MMAs)
    }
    // ---- epilogue 2: ELU(D + b2) . w3 + b3, zero diagonal / flag mask (ScoreNetwork_A.py:143-148), mirrored store
    float dot = 0.f;
    {
      constexpr int NB8 = CH / 8;
      constexpr int GB = NB8 % 3 == 0 ? 3 : (NB8 % 2 == 0 ? 2 : 1);
      float dot2 = 0.f;
#pragma unroll
      for (int b0 = 0; b0 < NB8; b0 += GB) {
        uint32_t dr[GB][8];
#pragma unroll
        for (int u = 0; u < GB; ++u) tmem_ld8_nowait(tlane + colD2 + half * CH + 8 * (b0 + u), dr[u]);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < GB; ++u) {
          const int col = half * CH + 8 * (b0 + u);
#pragma unroll
          for (int q = 0; q < 8; q += 2) {
            dot = fmaf(elu_f(__uint_as_float(dr[u][q]) + b2s[col + q]), w3s[col + q], dot);
            dot2 = fmaf(elu_f(__uint_as_float(dr[u][q + 1]) + b2s[col + q + 1]), w3s[col + q + 1], dot2);
          }
        }
      }
      dot += dot2;
    }
    if (half > 0) part[(half - 1) * 128 + row] = dot;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    ok = __syncthreads_and(ok);             // partial sums visible; every thread is done reading D before the next tile's MMA;
                                            // uniform loop exit.  part[] is rewritten only after the next tile's barrier
    if (half == 0) {
      if (out_b >= 0) {
        const int b = out_b, i = out_i, j = out_j;
        float acc = dot;
#pragma unroll
        for (int q = 0; q < NS - 1; ++q) acc += part[q * 128 + row];
        const float s = (acc + b3) * out_f;
        float* O = a.out + (int64_t)b * N * N;
        O[(int64_t)i * N + j] = s;
        O[(int64_t)j * N + i] = s;
      }
    }
  }
  if (!ok && tid == 0) atomicExch(err, 1);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tb) : "memory");
}

typedef void (*final_tc_fn)(FinalTriArgs, int*);
static final_tc_fn pick_final_tc(int fdim, int* np, int* k1p, int* ns) {
  const int H = 2 * fdim, NP = (H + 15) & ~15, K1P = (fdim + 7) & ~7;
  *np = NP;
  *k1p = K1P;
  *ns = 2;
  if (NP == 96 && K1P == 48) return final_edge_tc_kernel<96, 48, 4, false, 2>;  // fdim 46 (ZINC250k: 2 + 8*5 + 4); NS = 4 measured 3 % slower
  if (NP == 80 && K1P == 40) return final_edge_tc_kernel<80, 40, 4, false, 2>;     // fdim 38 (community_small, ego_small)
  if (NP == 48 && K1P == 24) return final_edge_tc_kernel<48, 24, 4, false, 2>;     // fdim 22 (QM9)
  if (NP == 112 && K1P == 56) return final_edge_tc_kernel<112, 56, 4, false, 2>;   // fdim 54 (ENZYMES, grid)
  return nullptr;
}

// node mode: the X networks whose hidden width fits tensor memory (3 NP <= 512 columns)
static final_tc_fn pick_final_tc_node(int fdim, int foutp, int* np, int* k1p, int* fop, int* ns) {
  const int H = 2 * fdim, NP = (H + 15) & ~15, K1P = (fdim + 7) & ~7;
  *np = NP;
  *k1p = K1P;
  *fop = foutp;
  *ns = 2;
  if (NP == 96 && K1P == 48 && foutp == 12) return final_edge_tc_kernel<96, 48, 12, true, 2>;   // ZINC250k: fdim 9 + 2*16, F = 9
  if (NP == 80 && K1P == 40 && foutp == 4) return final_edge_tc_kernel<80, 40, 4, true, 2>;     // QM9: fdim 4 + 2*16, F = 4
  return nullptr;
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
    if (L.hid != 16 || L.c_in > 8 || L.c_out > 8 || (L.c_out & 3)) return false;
  }
  return true;
}

static void fill_net(const NetPlan& p, const float* blob, FNet* f) {
  memset(f, 0, sizeof(*f));
  f->type = p.type; f->depth = p.depth; f->F = p.F; f->nhid = p.nhid; f->heads = p.heads;
  f->fdim = p.fdim; f->fout = p.fout; f->c_init = p.c_init; f->Hp = p.Hp; f->foutp = p.foutp;
  for (int l = 0; l < p.depth && l < FUSED_MAX_LAYERS; ++l) {
    if (p.type == GDSS_NET_X) {
      f->gw[l] = blob + p.gw[l];
      f->gb[l] = blob + p.gb[l];
      continue;
    }
    const LayerPlan& L = p.layers[l];
    FLayer& o = f->L[l];
    const bool last = l == p.depth - 1;
    o.c_in = L.c_in; o.c_out = L.c_out; o.f_in = L.f_in; o.f_in_p = L.f_in_p; o.attn = L.attn; o.nlin = L.nlin; o.qkvw = L.qkvw;
    o.need_edge = (p.type == GDSS_NET_A || !last) ? 1 : 0;   // X_GMH: the last layer's adj_out is never read
    o.need_node = ((p.type != GDSS_NET_A || !last) && L.node_live) ? 1 : 0;   // A: the last layer's x_out is never read
    o.edge_const = (o.need_edge && L.edge_const) ? 1 : 0;
    o.qk_live = (o.need_edge && !o.edge_const) ? L.qk_live : 0u;
    o.v_live = o.need_node ? L.v_live : 0u;
    o.nact = o.nqk = 0;
    for (int c = 0; c < L.c_in && c < 8; ++c) {
      if (((o.qk_live | o.v_live) >> c) & 1u) o.act[o.nact++] = (signed char)c;
      if ((o.qk_live >> c) & 1u) o.qkc[o.nqk++] = (signed char)c;
    }
    for (int c = 0; c < 8; ++c) o.cst[c] = c < L.c_out ? L.edge_cst[c] : 0.f;
    o.vk_live = 0ull;
    for (int c = 0; c < L.c_in && c < 8; ++c)
      if ((o.v_live >> c) & 1u)
        for (int k = c * L.nhid / 8; k < (c + 1) * L.nhid / 8 && k < 64; ++k) o.vk_live |= 1ull << k;
    o.blk_q = blob + L.wqkv;
    o.bq_o = (int)(L.bqkv - L.wqkv);
    o.q_floats = (int)(L.ew[0] - L.wqkv);                        // wqkv + bqkv (both padded to 4 floats)
    o.blk_m = blob + L.ew[0];
    for (int i = 0; i < L.nlin && i < 3; ++i) { o.ew_o[i] = (int)(L.ew[i] - L.ew[0]); o.eb_o[i] = (int)(L.eb[i] - L.ew[0]); }
    o.nw0_o = (int)(L.nw[0] - L.ew[0]); o.nb0_o = (int)(L.nb[0] - L.ew[0]);
    o.nw1_o = (int)(L.nw[1] - L.ew[0]); o.nb1_o = (int)(L.nb[1] - L.ew[0]);
    o.m_floats = (int)(L.nb[1] - L.ew[0]) + ((L.nhid + 3) & ~3);
  }
  for (int i = 0; i < 3; ++i) { f->fw[i] = blob + p.fw[i]; f->fb[i] = blob + p.fb[i]; }
}

// shared-memory plan for graphs of at most NB nodes; returns bytes (or -1 when the regions cannot host the
// dense adjacency alias)
static int64_t plan_smem(const gdss_ctx* ctx, int NB, FSmem* s, bool wsm = true) {
  const NetPlan* nets[2] = {ctx->has_a ? &ctx->pa : nullptr, ctx->has_x ? &ctx->px : nullptr};
  int Cmax = 1, hmax = 8, amax = 0, fmax = (ctx->F + 7) & ~7, fdx = 4, Hx = 0, cinit = 1;
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
  if (fmax < ((hmax + 7) & ~7)) fmax = (hmax + 7) & ~7;
  const int NBP = (NB + 3) & ~3;
  // plane stride: the P pixels + two constant slots (PM-2 holds 1.0 = the self loop, PM-1 holds 0.0 = padding; both are
  // addressed through the triangle-index table), rounded up to 8 (mod 32) so that the four planes an mma A fragment
  // gathers from fall into disjoint banks
  int PM = NB * (NB - 1) / 2 + 2;
  while ((PM & 31) != 8) ++PM;
  s->NBP = NBP; s->PM = PM;
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
  s->xs = take(NB * s->XSS);
  s->h1 = take(NB * 16);
  int wqf = 4, wmf = 4;
  for (int k = 0; k < 2; ++k) {
    const NetPlan* p = nets[k];
    if (!p || p->type == GDSS_NET_X) continue;
    for (int l = 0; l < p->depth; ++l) {
      const LayerPlan& L = p->layers[l];
      const int q = (int)(L.ew[0] - L.wqkv), m = (int)(L.nb[1] - L.ew[0]) + ((L.nhid + 3) & ~3);
      wqf = wqf > q ? wqf : q;
      wmf = wmf > m ? wmf : m;
    }
  }
  s->wq = take(wsm ? wqf : 4);
  s->wm = take(wsm ? wmf : 4);
  s->p0 = take(cinit * s->PM);
  s->pa = take(Cmax * s->PM);
  s->RS = NB * fmax > s->PM ? NB * fmax : s->PM;     // per channel: A.x rows [n][f rounded up to 8], then attention map [P]
  while ((s->RS & 31) != 8) ++s->RS;
  int r1 = Cmax * s->RS;
  if (r1 < NB * s->LD) r1 = NB * s->LD;              // dense A while the power stack is built
  int qk = Cmax * NB * s->QKS;
  if (r1 + qk < 2 * NB * Hx) qk = 2 * NB * Hx - r1;  // [R1 | QK] hosts the two hidden buffers of the final node MLP
  s->r1 = take(r1);
  s->qk = take(qk);
  s->r2 = take(NB * (Cmax * hmax + 4));            // V rows [i][c][h], row stride c_in * nhid + 4
  s->TC = 8 * ((NB + 7) / 8);
  s->tab = take((16 * ((NB + 15) / 16) * s->TC + 1) / 2);      // uint16 entries
  s->total = cur;
  return (int64_t)cur * 4;
}

// size buckets of the per-graph kernel (its shared memory is sized by the largest graph of a launch):
// graphs of at most 23 / 32 / N nodes; returns the number of buckets and their upper bounds
static int size_buckets(int N, int nb[3]) {
  int k = 0;
  if (N > 23) nb[k++] = 23;
  if (N > 32) nb[k++] = 32;
  nb[k++] = N;
  return k;
}

bool fused_supported(const gdss_ctx* ctx) {
  if (ctx->N > 64 || ctx->N < 2) return false;
  if (ctx->has_a && !net_fusable(ctx->pa)) return false;
  if (ctx->has_x && !net_fusable(ctx->px)) return false;
  FSmem s;
  const int64_t bytes = plan_smem(ctx, ctx->N, &s);
  if (bytes > 0 && bytes <= ctx->smem_optin) return true;
  const int64_t bytes_l2 = plan_smem(ctx, ctx->N, &s, false);       // layer weights read through L2 instead of staged
  return bytes_l2 > 0 && bytes_l2 <= ctx->smem_optin;
}

int64_t fused_query(const gdss_ctx* ctx, int what) {
  const bool ok = ctx->N <= 64 && ctx->N >= 2 && (!ctx->has_a || net_fusable(ctx->pa)) && (!ctx->has_x || net_fusable(ctx->px));
  if (what == 0) return ok ? 1 : 0;
  if (!ok) return -1;
  FSmem s;
  if (what == 1) return plan_smem(ctx, ctx->N, &s);
  // what = 1000 + 2 * NB + wsm: the plan of a size bucket of at most NB nodes, with / without staged weights
  const int nb = (what - 1000) >> 1;
  if (nb < 2 || nb > ctx->N) return -1;
  return plan_smem(ctx, nb, &s, (what & 1) != 0);
}

int launch_tri_setup(const gdss_ctx* ctx, const int* neff, int B, char* ws, const Workspace& w, cudaStream_t st) {
  int* poff = reinterpret_cast<int*>(ws + w.off[R_POFF]);
  int* meta = reinterpret_cast<int*>(ws + w.off[R_META]);
  int* gmap = reinterpret_cast<int*>(ws + w.off[R_GMAP]);
  int* order = reinterpret_cast<int*>(ws + w.off[R_ORDER]);
  int nb[3];
  const int k = size_buckets(ctx->N, nb);
  // scan kernel buckets: n <= nb0 -> 0, n <= nb1 -> 1, else 2 (unused buckets get bounds that never match)
  const int b0 = nb[0], b1 = k > 1 ? nb[1] : nb[0];
  int* noff = ctx->fused ? reinterpret_cast<int*>(ws + w.off[R_NOFF]) : nullptr;
  int* nmap = ctx->fused ? reinterpret_cast<int*>(ws + w.off[R_NMAP]) : nullptr;
  tri_scan_kernel<<<1, 1024, 0, st>>>(neff, B, b0, b1, poff, meta, gmap, order, noff, nmap);
  launched(K_MISC, st);
  LAUNCH_CHECK();
  tri_map_kernel<<<B, 64, 0, st>>>(neff, poff, gmap, ctx->fused ? 6 : 9, noff, nmap);
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
  fa.noff = reinterpret_cast<const int*>(ws + w.off[R_NOFF]);
  fa.pstride = w.pstride;
  fa.N = N; fa.F = ctx->F;
  // X network's final MLP: batched tensor-core kernel when its shape has an instance, else inside the per-graph kernel
  int xnp = 0, xk1p = 0, xfop = 0, xns = 2;
  final_tc_fn xfn = (want_x && ctx->use_tc && w.nstride > 0) ? pick_final_tc_node(ctx->px.fdim, ctx->px.foutp, &xnp, &xk1p, &xfop, &xns) : nullptr;
  const size_t xsm = sizeof(float) * (2 * (size_t)(xk1p * xnp + xnp * xnp) + 2 * xnp + (size_t)xnp * xfop + 128 * (size_t)xfop * (xns - 1)) + 256;
  if (xfn && (int)xsm > ctx->smem_optin) xfn = nullptr;
  if (xfn) {
    fa.xst = reinterpret_cast<float*>(ws + w.off[R_XS]);
    fa.nstride = w.nstride;
  }
  fa.use_mma = ctx->use_mma ? 1 : 0;
  fa.use_tc_edge = ctx->use_tc_edge ? 1 : 0;
  fa.tce_slots = EDGE_SLOTS;
  if (const char* ts = getenv("GDSS_TCE_SLOTS")) { const int v = atoi(ts); if (v >= 1 && v <= EDGE_SLOTS) fa.tce_slots = v; }
  fa.prof = ctx->phase_prof ? reinterpret_cast<unsigned long long*>(ws + w.off[R_META] + 128) : nullptr;
  fa.order = reinterpret_cast<const int*>(ws + w.off[R_ORDER]);
  fa.meta = reinterpret_cast<const int*>(ws + w.off[R_META]);
  int nb[3];
  const int nbk = size_buckets(N, nb);
  // The size buckets are independent: they run as parallel branches (the largest graphs first, on the caller's stream,
  // the others on the context's side streams; fork / join by events, also when the caller is capturing a CUDA graph), so
  // a bucket that does not fill the GPU (few large graphs, small shards at 8 GPUs) overlaps with the others.
  const bool fork = nbk > 1 && ctx->side[0] && ctx->side[1];
  if (fork) {
    cudaError_t fe = cudaEventRecord(ctx->ev_fork, st);
    for (int k = 0; k < nbk - 1 && fe == cudaSuccess; ++k) fe = cudaStreamWaitEvent(ctx->side[k], ctx->ev_fork, 0);
    if (fe != cudaSuccess) return (int)fe;
  }
  const bool tce = ctx->use_tc_edge && ctx->use_mma;
  for (int kk = 0; kk < nbk; ++kk) {
    const int k = nbk - 1 - kk;                         // largest bucket first
    cudaStream_t ks = (fork && kk < nbk - 1) ? ctx->side[kk] : st;     // the smallest graphs on the caller's (lowest-priority) stream
    // one launch per size bucket: B CTAs, those beyond the bucket's population exit at once (the counts live
    // on the device; no host synchronisation)
    fa.bucket = k;
    fa.mb1 = nb[k] <= 16 ? 1 : 0;
    // Preferred: layer weights staged in shared memory by TMA with two CTAs of 256 threads per SM; if only the
    // weight-less plan fits twice, read the weights through L2 instead; the largest graphs run one CTA of 512.
    // (three CTAs of 85 registers per SM with the weights read through L2 measured 5 % slower than two with staged weights;
    //  re-measured on the final build for the n <= 23 bucket: 80 registers, 388 B of spills, 6 % slower)
    const int limit2 = (227 * 1024) / 2 - 1024;
    int64_t smem = plan_smem(ctx, nb[k], &fa.sl, true);
    cudaError_t e;
    if (nb[k] <= 12) {
      // tiny graphs (QM9): 128-thread CTAs; without the staged weights (38 KB) four of them fit an SM (register bound)
      // and the weights of all layers stay in the then large L1 -- measured +18 % on C2
      smem = plan_smem(ctx, nb[k], &fa.sl, false);
      // (64 / 96 / 256 threads per CTA measured 20 % / 12 % / 31 % slower than 128 on C2)
      // (register caps for 5 / 6 / 8 CTAs per SM -- 96 / 80 / 64 registers, 0.3 .. 0.7 KB of spills -- measured 1 % / 6 % / 11 %
      //  slower than the 128-register build at 4 CTAs per SM: 10.46 k -> 10.35 k / 9.88 k / 9.27 k graphs/s on C2)
      if (int rc_ = ensure_smem((const void*)fused_layers_kernel<256, false, 2>, smem)) return rc_;
      fused_layers_kernel<256, false, 2><<<B, 128, smem, ks>>>(fa);
    } else if (smem <= limit2) {
      void (*kf)(const FusedArgs) = tce ? fused_layers_kernel<256, true, 2, true> : fused_layers_kernel<256, true, 2>;
      if (int rc_ = ensure_smem((const void*)kf, smem)) return rc_;
      kf<<<B, 256, smem, ks>>>(fa);
    } else if (plan_smem(ctx, nb[k], &fa.sl, false) <= limit2) {   // (one CTA of 512 with staged weights measured 20 % slower here)
      smem = plan_smem(ctx, nb[k], &fa.sl, false);
      void (*kf)(const FusedArgs) = tce ? fused_layers_kernel<256, false, 2, true> : fused_layers_kernel<256, false, 2>;
      if (int rc_ = ensure_smem((const void*)kf, smem)) return rc_;
      kf<<<B, 256, smem, ks>>>(fa);
    } else if (smem > 0 && smem <= ctx->smem_optin) {
      smem = plan_smem(ctx, nb[k], &fa.sl, true);
      void (*kf)(const FusedArgs) = tce ? fused_layers_kernel<512, true, 1, true> : fused_layers_kernel<512, true, 1>;
      if (int rc_ = ensure_smem((const void*)kf, smem)) return rc_;
      kf<<<B, 512, smem, ks>>>(fa);
    } else {
      // wide layers (nhid = adim = 32, 8 channels: community_small's 96 KB of Q|K|V weights per layer): the staged plan
      // does not fit an SM at all; one CTA of 256 per SM with the weights read through L1 / L2
      smem = plan_smem(ctx, nb[k], &fa.sl, false);
      if (smem <= 0 || smem > ctx->smem_optin) return GDSS_E_UNSUPPORTED;
      void (*kf)(const FusedArgs) = tce ? fused_layers_kernel<256, false, 2, true> : fused_layers_kernel<256, false, 2>;
      if (int rc_ = ensure_smem((const void*)kf, smem)) return rc_;
      kf<<<B, 256, smem, ks>>>(fa);
    }
    count_launch();
    if (!fork) prof_tick(K_FUSED, st);
    LAUNCH_CHECK();
    if (fork && kk < nbk - 1) {
      cudaError_t je = cudaEventRecord(ctx->ev_join[kk], ks);
      if (je != cudaSuccess) return (int)je;
    }
  }
  if (fork) {
    for (int k = 0; k < nbk - 1; ++k)
      if (cudaStreamWaitEvent(st, ctx->ev_join[k], 0) != cudaSuccess) return (int)cudaGetLastError();
    prof_tick(K_FUSED, st);                             // one interval for the parallel branches (after the join)
  }
  cudaError_t e;
  bool x_side = false;
  if (xfn) {
    const NetPlan& p = ctx->px;
    FinalTriArgs ta;
    memset(&ta, 0, sizeof(ta));
    ta.stackT = fa.xst; ta.pstride = w.nstride;
    ta.meta = reinterpret_cast<const int*>(ws + w.off[R_META]);
    ta.w1 = ctx->blob_x + p.fw[0]; ta.b1 = ctx->blob_x + p.fb[0]; ta.w2 = ctx->blob_x + p.fw[1]; ta.b2 = ctx->blob_x + p.fb[1];
    ta.w3 = ctx->blob_x + p.fw[2]; ta.b3 = ctx->blob_x + p.fb[2];
    ta.out = score_x; ta.flags = flags; ta.N = N; ta.fdim = p.fdim; ta.H = 2 * p.fdim; ta.Hs = p.Hp;
    ta.neff = neff; ta.B = B; ta.F = ctx->F; ta.foutp = p.foutp;
    ta.nmap = reinterpret_cast<const int*>(ws + w.off[R_NMAP]);
    if (int rc_ = ensure_smem((const void*)xfn, xsm)) return rc_;
    const int64_t tiles = ((int64_t)B * N + 127) / 128;
    int grid = ctx->num_sms > 0 ? ctx->num_sms : 148;
    if ((int64_t)grid > tiles) grid = (int)tiles;
    int* errflag = reinterpret_cast<int*>(ws + w.off[R_META]) + 16;
    // The X network's final MLP and ScoreNetworkA.final are independent: the former runs on a side stream next to the latter
    // (each is one persistent wave of at most num_sms CTAs whose prologue / tail the other fills; matters on the small shards of
    // an 8-GPU run).  Serial under the per-kernel profiler, which times on the caller's stream.
    x_side = fork && want_a && !prof_active();
    cudaStream_t xs_ = st;
    if (x_side) {
      cudaError_t fe = cudaEventRecord(ctx->ev_fork, st);
      if (fe == cudaSuccess) fe = cudaStreamWaitEvent(ctx->side[0], ctx->ev_fork, 0);
      if (fe != cudaSuccess) return (int)fe;
      xs_ = ctx->side[0];
    }
    xfn<<<grid, 128 * xns, xsm, xs_>>>(ta, errflag);
    launched(K_FINAL_NODE, st);
    LAUNCH_CHECK();
    if (x_side) {
      cudaError_t je = cudaEventRecord(ctx->ev_join[0], xs_);
      if (je != cudaSuccess) return (int)je;
    }
  }
  if (want_a) {
    const NetPlan& p = ctx->pa;
    FinalTriArgs ta;
    memset(&ta, 0, sizeof(ta));
    ta.stackT = fa.stackT; ta.pstride = w.pstride;
    ta.gmap = reinterpret_cast<const int*>(ws + w.off[R_GMAP]);
    ta.meta = reinterpret_cast<const int*>(ws + w.off[R_META]);
    ta.w1 = ctx->blob_a + p.fw[0]; ta.b1 = ctx->blob_a + p.fb[0]; ta.w2 = ctx->blob_a + p.fw[1]; ta.b2 = ctx->blob_a + p.fb[1];
    ta.w3 = ctx->blob_a + p.fw[2]; ta.b3 = ctx->blob_a + p.fb[2];
    ta.out = score_adj; ta.flags = flags; ta.N = N; ta.fdim = p.fdim; ta.H = 2 * p.fdim; ta.Hs = p.Hp;
    ta.fb = 6;
    const int64_t max_tiles = ((int64_t)B * (N * (N - 1) / 2) + 127) / 128;
    int np = 0, k1p = 0, tns = 2;
    final_tc_fn tfn = ctx->use_tc ? pick_final_tc(p.fdim, &np, &k1p, &tns) : nullptr;
    if (tfn && max_tiles > 0) {
      const size_t tsm = sizeof(float) * (2 * (size_t)(k1p * np + np * np) + 3 * np + 128 * (size_t)(tns - 1)) + 256;
      if ((int)tsm <= ctx->smem_optin) {
        if (int rc_ = ensure_smem((const void*)tfn, tsm)) return rc_;
        int grid = ctx->num_sms > 0 ? ctx->num_sms : 148;
        if ((int64_t)grid > max_tiles) grid = (int)max_tiles;
        int* errflag = reinterpret_cast<int*>(ws + w.off[R_META]) + 16;
        tfn<<<grid, 128 * tns, tsm, st>>>(ta, errflag);
        launched(K_FINAL_EDGE, st);
        LAUNCH_CHECK();
        if (x_side && cudaStreamWaitEvent(st, ctx->ev_join[0], 0) != cudaSuccess) return (int)cudaGetLastError();
        return 0;
      }
    }
    int HP;
    final_tri_fn fn = pick_final_tri(ta.H, &HP);
    if (!fn) return GDSS_E_UNSUPPORTED;
    int64_t sizeA = (int64_t)p.fdim * 128 + (int64_t)p.fdim * HP;
    if (sizeA < (int64_t)ta.H * HP) sizeA = (int64_t)ta.H * HP;
    const size_t fsm = sizeof(float) * (sizeA + (size_t)HP * 128 + 3 * HP);
    if ((int)fsm > ctx->smem_optin) return GDSS_E_UNSUPPORTED;
    if (fsm > 48 * 1024) {
      if (int rc_ = ensure_smem((const void*)fn, fsm)) return rc_;
    }
    if (max_tiles > 0) {
      fn<<<(unsigned)max_tiles, 256, fsm, st>>>(ta);
      launched(K_FINAL_EDGE, st);
      LAUNCH_CHECK();
    }
  }
  if (x_side && cudaStreamWaitEvent(st, ctx->ev_join[0], 0) != cudaSuccess) return (int)cudaGetLastError();
  return 0;
}

// ---- general path (N > 64): ScoreNetworkA.final on the tensor-core kernel, reading the dense channel stack
static size_t final_tc_smem(int np, int k1p, int ns) {
  return sizeof(float) * (2 * (size_t)(k1p * np + np * np) + 3 * np + 128 * (size_t)(ns - 1)) + 256;
}

bool final_tc_dense_ok(const gdss_ctx* ctx, int B) {
  if (ctx->fused || !ctx->has_a || !ctx->use_tc || ctx->N > 511 || B >= 8192) return false;
  if ((int64_t)B * (ctx->N * (ctx->N - 1) / 2) >= (int64_t)1 << 30) return false;
  int np, k1p, ns;
  if (!pick_final_tc(ctx->pa.fdim, &np, &k1p, &ns)) return false;
  return (int)final_tc_smem(np, k1p, ns) <= ctx->smem_optin;
}

int launch_final_tc_dense(const gdss_ctx* ctx, const float* stack, int64_t s_bs, const float* flags, float* score_adj, int B,
                          char* ws, const Workspace& w, cudaStream_t st) {
  const NetPlan& p = ctx->pa;
  const int N = ctx->N;
  FinalTriArgs ta;
  memset(&ta, 0, sizeof(ta));
  ta.stackT = stack; ta.pstride = 0; ta.dense_bs = s_bs; ta.fb = 9;
  ta.gmap = reinterpret_cast<const int*>(ws + w.off[R_GMAP]);
  ta.meta = reinterpret_cast<const int*>(ws + w.off[R_META]);
  ta.w1 = ctx->blob_a + p.fw[0]; ta.b1 = ctx->blob_a + p.fb[0]; ta.w2 = ctx->blob_a + p.fw[1]; ta.b2 = ctx->blob_a + p.fb[1];
  ta.w3 = ctx->blob_a + p.fw[2]; ta.b3 = ctx->blob_a + p.fb[2];
  ta.out = score_adj; ta.flags = flags; ta.N = N; ta.fdim = p.fdim; ta.H = 2 * p.fdim; ta.Hs = p.Hp;
  int np = 0, k1p = 0, tns = 2;
  final_tc_fn tfn = pick_final_tc(p.fdim, &np, &k1p, &tns);
  if (!tfn) return GDSS_E_UNSUPPORTED;
  const size_t tsm = final_tc_smem(np, k1p, tns);
  if (int rc_ = ensure_smem((const void*)tfn, tsm)) return rc_;
  const int64_t max_tiles = ((int64_t)B * (N * (N - 1) / 2) + 127) / 128;
  if (max_tiles <= 0) return 0;
  int grid = ctx->num_sms > 0 ? ctx->num_sms : 148;
  if ((int64_t)grid > max_tiles) grid = (int)max_tiles;
  int* errflag = reinterpret_cast<int*>(ws + w.off[R_META]) + 16;
  tfn<<<grid, 128 * tns, tsm, st>>>(ta, errflag);
  launched(K_FINAL_EDGE, st);
  LAUNCH_CHECK();
  return 0;
}

}  // namespace gdss
