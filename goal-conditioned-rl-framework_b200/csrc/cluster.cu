// Cluster-fused DDPG update kernels (small-batch regime, hidden width 64..256).
//
// Same math as fused.cu (critic phase: reference src/agent.py:1302-1329, actor phase :1288-1294) but
// every hidden layer is split over the 8 CTAs of a thread-block cluster by OUTPUT column:
//   * a cluster owns 32 batch rows; CTA `rank` owns columns [rank*w, (rank+1)*w) of every layer, w = H/8;
//   * each CTA streams only ITS column slice of the weight matrix (K x w floats, double-buffered
//     cp.async prefetch one layer ahead), so a layer's weights cross the L2 -> SM fabric once per
//     cluster instead of once per CTA (8x less traffic than the row-slab kernels);
//   * the [w x 32] output slice is epilogued (bias + LeakyReLU, or LeakyReLU') and written into the
//     next-layer input buffer of all 8 CTAs through distributed shared memory, then one
//     barrier.cluster separates the layers.
// The skinny heads (<= 4 outputs), the Bellman target / loss and the action gradient are computed
// redundantly by every CTA from the replicated activations, so they need no exchange.
// Fixed summation order everywhere (K interleaved over thread groups, then a fixed-order combine):
// deterministic run to run.
#include <cooperative_groups.h>

#include <algorithm>

#include "mlp.cuh"

namespace cg = cooperative_groups;

namespace gcrl {
namespace {

constexpr int CS = 8;      // CTAs per cluster (portable maximum)
constexpr int CT = 512;    // threads per CTA
constexpr int RC = 32;     // batch rows per cluster (256 rows = 8 clusters = one wave; at most 15 clusters of 8
                           // are co-resident on the 148 SMs)
constexpr int kMaxSteps = 4 * kFusedMaxL;

struct MatStep {
  const float *M;   // [K][ld]: Wt (forward) or W (input gradient); this CTA uses columns rank*w .. +w
  int ld, K;
};

struct ClusterCommon {
  MatStep steps[kMaxSteps];
  int nsteps;
  int B, D, A, H, L, ldh, ldc, w;
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  const uint32_t d = uint32_t(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct Ctx {
  cg::cluster_group cluster;
  int rank, w, tid;
  float *wb[2];        // weight slice double buffer [K][w]
  float *red;          // [CT * 16] K-split partials
  float *ys;           // [w][RC] epilogued output slice of this CTA
  const ClusterCommon *cc;
  int next_step;       // next step whose weights have NOT been requested yet
};

// request the weight slice of step s into wb[s & 1]
__device__ __forceinline__ void prefetch_step(Ctx &c, int s) {
  if (s < c.cc->nsteps) {
    const MatStep &st = c.cc->steps[s];
    const int w4 = c.w >> 2;
    float *dst = c.wb[s & 1];
    const float *src = st.M + c.rank * c.w;
    for (int p = c.tid; p < st.K * w4; p += CT) {
      const int row = p / w4, c4 = p - row * w4;
      cp_async16(dst + row * c.w + c4 * 4, src + size_t(row) * st.ld + c4 * 4);
    }
  }
  cp_async_commit();
}

enum : int { EP_BIAS_LEAKY = 0, EP_DLEAKY = 1 };

// One layer step.  xT: full input [K][RC] (replicated in every CTA); result slice -> c.ys [w][RC],
// also broadcast into rows [rank*w, +w) of `xnext` in all CTAs.  Ends with a cluster barrier.
//   EP_BIAS_LEAKY: leaky(acc + bias[rank*w + j]);  EP_DLEAKY: acc * leaky'(ref[j][r]) (ref = kept slice)
template <int EPI>
__device__ __forceinline__ void cluster_layer(Ctx &c, int s, const float *xT, const float *__restrict__ bias,
                                              const float *ref, float *xnext) {
  const MatStep &st = c.cc->steps[s];
  const int w = c.w, w4 = w >> 2, tpk = w4 * (RC / 4), nsl = CT / tpk;
  const int ks = c.tid / tpk, within = c.tid - ks * tpk;
  const int cgp = within % w4, rg = within / w4;
  cp_async_wait_all();           // this step's weights (requested one step ago)
  __syncthreads();
  prefetch_step(c, s + 1);       // overlaps with the math below
  const float *wbuf = c.wb[s & 1];
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int kr = ks; kr < st.K; kr += nsl) {
    const float4 wv = *reinterpret_cast<const float4 *>(wbuf + kr * w + cgp * 4);
    const float4 xv = *reinterpret_cast<const float4 *>(xT + kr * RC + rg * 4);
    const float wa[4] = {wv.x, wv.y, wv.z, wv.w}, xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wa[i], xa[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4 *>(c.red + size_t(ks) * (w * RC) + (cgp * 4 + i) * RC + rg * 4) =
        make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  __syncthreads();
  for (int e = c.tid; e < w * RC; e += CT) {
    float v = c.red[e];
    for (int k2 = 1; k2 < nsl; ++k2) v += c.red[size_t(k2) * (w * RC) + e];
    if (EPI == EP_BIAS_LEAKY) {
      v += bias[c.rank * w + e / RC];
      v = v > 0.f ? v : v * kLeakySlope;
    } else {
      v = ref[e] > 0.f ? v : v * kLeakySlope;
    }
    c.ys[e] = v;
  }
  __syncthreads();
  // broadcast the slice into every CTA's next-layer input (distributed shared memory)
  const int pieces = (w * RC) >> 2;
  for (int idx = c.tid; idx < CS * pieces; idx += CT) {
    const int d = idx / pieces, p = idx - d * pieces;
    float *remote = c.cluster.map_shared_rank(xnext, d);
    *reinterpret_cast<float4 *>(remote + c.rank * w * RC + p * 4) = *reinterpret_cast<const float4 *>(c.ys + p * 4);
  }
  c.cluster.sync();
}

// out[j][r] = f(sum_k hT[k][r] Wh[j][k] + bh[j]), j < nout <= 4: one warp per (j, r); every CTA, redundantly
__device__ __forceinline__ void head_full(const float *hT, int K, const float *__restrict__ Wh, int ldw,
                                          const float *__restrict__ bh, int nout, bool tanh_out, float *out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int p = warp; p < nout * RC; p += CT / 32) {
    const int j = p / RC, r = p - j * RC;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(hT[k * RC + r], __ldg(Wh + size_t(j) * ldw + k), acc);
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) {
      const float v = acc + bh[j];
      out[p] = tanh_out ? tanhf(v) : v;
    }
  }
  __syncthreads();
}

// this CTA's [w][RC] slice -> global [row0 + r][ld] columns rank*w .. (rows >= B skipped)
__device__ __forceinline__ void store_slice(const Ctx &c, const float *ys, float *__restrict__ out, int ld, int row0,
                                            int B) {
  for (int e = c.tid; e < c.w * RC; e += CT) {
    const int r = e / c.w, j = e - r * c.w;
    if (row0 + r < B) out[size_t(row0 + r) * ld + c.rank * c.w + j] = ys[j * RC + r];
  }
}
__device__ __forceinline__ void copy_slice(const Ctx &c, const float *src, float *dst) {
  for (int e = c.tid; e < c.w * RC; e += CT) dst[e] = src[e];
}

struct Carve {
  float *xns, *xsa, *h0, *h1, *wb0, *wb1, *red, *ys, *small, *keep[2 * kFusedMaxL];
};
__device__ __forceinline__ Carve carve(float *base, int KinP, int H, int L, int w) {
  Carve p;
  p.xns = base; base += KinP * RC;
  p.xsa = base; base += KinP * RC;
  p.h0 = base; base += H * RC;
  p.h1 = base; base += H * RC;
  p.wb0 = base; base += H * w;
  p.wb1 = base; base += H * w;
  p.red = base; base += CT * 16;
  p.ys = base; base += w * RC;
  p.small = base; base += 16 * RC;
  for (int l = 0; l < 2 * L; ++l) { p.keep[l] = base; base += w * RC; }
  return p;
}
size_t cluster_smem_bytes(int D, int A, int H, int L) {
  const int KinP = (D + A + 3) & ~3, w = H / CS;
  const size_t f = size_t(2) * KinP * RC + size_t(2) * H * RC + size_t(2) * H * w + size_t(CT) * 16 + size_t(w) * RC +
                   16 * RC + size_t(2) * L * w * RC;
  return f * sizeof(float) + 16;
}

// ---------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(CT, 1)
cluster_critic_kernel(const __grid_constant__ ClusterCommon cc, FusedCriticArgs a) {
  extern __shared__ float4 csm4[];
  Ctx c{cg::this_cluster(), 0, cc.w, int(threadIdx.x), {nullptr, nullptr}, nullptr, nullptr, &cc, 0};
  c.rank = int(c.cluster.block_rank());
  const int D = cc.D, A = cc.A, H = cc.H, L = cc.L, B = cc.B, tid = c.tid;
  const int KinP = (D + A + 3) & ~3;
  Carve sp = carve(reinterpret_cast<float *>(csm4), KinP, H, L, cc.w);
  c.wb[0] = sp.wb0; c.wb[1] = sp.wb1; c.red = sp.red; c.ys = sp.ys;
  float *qn = sp.small, *q = qn + RC, *dzh = q + RC, *rr = dzh + RC, *dd = rr + RC, *anext = dd + RC;   // anext [A][RC]
  const int row0 = int(blockIdx.x / CS) * RC;

  prefetch_step(c, 0);
  for (int e = tid; e < KinP * RC; e += CT) {
    const int k = e / RC, r = e - k * RC, row = row0 + r;
    float vns = 0.f, vsa = 0.f;
    if (row < B) {
      if (k < D) { vns = a.ns[size_t(row) * D + k]; vsa = a.s[size_t(row) * D + k]; }
      else if (k < D + A) vsa = a.a[size_t(row) * A + (k - D)];
    }
    sp.xns[e] = vns;
    sp.xsa[e] = vsa;
  }
  if (tid < RC) {
    const int row = row0 + tid;
    rr[tid] = row < B ? a.r[row] : 0.f;
    dd[tid] = row < B ? a.d[row] : 0.f;
  }
  __syncthreads();
  if (c.rank == 0 && a.sa_out != nullptr) {
    for (int e = tid; e < a.ldc * RC; e += CT) {
      const int r = e / a.ldc, k = e - r * a.ldc;
      if (row0 + r < B) a.sa_out[size_t(row0 + r) * a.ldc + k] = k < KinP ? sp.xsa[k * RC + r] : 0.f;
    }
  }
  c.cluster.sync();      // every CTA of the cluster is running before the first remote store

  int s = 0;
  float *cur = sp.h0, *nxt = sp.h1;
  auto forward = [&](const FusedNet &n, const float *in, float *const *keep, float *const *gout) -> const float * {
    const float *x = in;
    for (int l = 0; l < L; ++l, ++s) {
      cluster_layer<EP_BIAS_LEAKY>(c, s, x, n.b[l], nullptr, nxt);
      if (keep) copy_slice(c, c.ys, keep[l]);
      if (gout) store_slice(c, c.ys, gout[l], a.ldh, row0, B);
      x = nxt;
      float *t = cur; cur = nxt; nxt = t;
    }
    return x;      // == cur
  };
  // 1. a' = target_actor(s')
  const float *h = forward(a.ta, sp.xns, nullptr, nullptr);
  head_full(h, H, a.ta.Wh, a.ta.ldwh, a.ta.bh, A, true, anext);
  for (int e = tid; e < A * RC; e += CT) sp.xns[D * RC + e] = anext[e];
  __syncthreads();
  // 2. q' = target_critic([s', a'])
  h = forward(a.tc, sp.xns, nullptr, nullptr);
  head_full(h, H, a.tc.Wh, a.tc.ldwh, a.tc.bh, 1, false, qn);
  // 3. q = critic([s, a]); own activation slices kept for the LeakyReLU' masks
  h = forward(a.c, sp.xsa, sp.keep, a.h_out);
  head_full(h, H, a.c.Wh, a.c.ldwh, a.c.bh, 1, false, q);
  // 4. Bellman target, loss, dL/dq (every CTA; rank 0 publishes)
  if (tid == 0) {
    float ls = 0.f, ts = 0.f, qs = 0.f;
    const float invB = 1.0f / float(B);
    for (int r = 0; r < RC; ++r) {
      const int row = row0 + r;
      float g = 0.f;
      if (row < B) {
        float y = rr[r] + a.gamma * (1.0f - dd[r]) * qn[r];
        if (a.clamp_y) y = fminf(fmaxf(y, a.y_lo), 0.0f);
        const float diff = q[r] - y;
        ls += diff * diff;
        ts += fabsf(y - q[r]);
        qs += q[r];
        g = 2.0f * diff * invB;
        if (c.rank == 0) {
          a.dzh_out[row] = g;
          if (a.y_out) a.y_out[row] = y;
          if (a.q_out) a.q_out[row] = q[r];
        }
      }
      dzh[r] = g;
    }
    if (c.rank == 0) {
      float *mp = a.metric_partials + size_t(blockIdx.x / CS) * 4;
      mp[0] = ls; mp[1] = ts; mp[2] = qs; mp[3] = 0.f;
    }
  }
  __syncthreads();
  // 5. backward: dz of the last hidden layer (full vector, every CTA), then layer by layer
  for (int e = tid; e < H * RC; e += CT) {
    const int k = e / RC, r = e - k * RC;
    const float g = dzh[r] * __ldg(a.c.Wh + k);
    nxt[e] = cur[e] > 0.f ? g : g * kLeakySlope;
  }
  c.cluster.sync();      // peers are done reading `cur` before the next layer's remote stores land in it
  { float *t = cur; cur = nxt; nxt = t; }
  store_slice(c, cur + c.rank * cc.w * RC, a.dz_out[L - 1], a.ldh, row0, B);
  for (int l = L - 1; l >= 1; --l, ++s) {
    cluster_layer<EP_DLEAKY>(c, s, cur, nullptr, sp.keep[l - 1], nxt);
    store_slice(c, c.ys, a.dz_out[l - 1], a.ldh, row0, B);
    float *t = cur; cur = nxt; nxt = t;
  }
  cp_async_wait_all();
  c.cluster.sync();      // no CTA exits while a peer may still address its shared memory
}

// ---------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(CT, 1)
cluster_actor_kernel(const __grid_constant__ ClusterCommon cc, FusedActorArgs a) {
  extern __shared__ float4 csm4[];
  Ctx c{cg::this_cluster(), 0, cc.w, int(threadIdx.x), {nullptr, nullptr}, nullptr, nullptr, &cc, 0};
  c.rank = int(c.cluster.block_rank());
  const int D = cc.D, A = cc.A, H = cc.H, L = cc.L, B = cc.B, tid = c.tid;
  const int KinP = (D + A + 3) & ~3;
  Carve sp = carve(reinterpret_cast<float *>(csm4), KinP, H, L, cc.w);
  c.wb[0] = sp.wb0; c.wb[1] = sp.wb1; c.red = sp.red; c.ys = sp.ys;
  float *q = sp.small, *act = q + RC, *da = act + 4 * RC;   // act, da: [A][RC]
  float *const *keep_a = sp.keep, *const *keep_c = sp.keep + L;
  const int row0 = int(blockIdx.x / CS) * RC;

  prefetch_step(c, 0);
  for (int e = tid; e < KinP * RC; e += CT) {
    const int k = e / RC, r = e - k * RC, row = row0 + r;
    sp.xsa[e] = (row < B && k < D) ? a.s[size_t(row) * D + k] : 0.f;
  }
  __syncthreads();
  c.cluster.sync();

  int s = 0;
  float *cur = sp.h0, *nxt = sp.h1;
  auto forward = [&](const FusedNet &n, const float *in, float *const *keep, float *const *gout) -> const float * {
    const float *x = in;
    for (int l = 0; l < L; ++l, ++s) {
      cluster_layer<EP_BIAS_LEAKY>(c, s, x, n.b[l], nullptr, nxt);
      copy_slice(c, c.ys, keep[l]);
      if (gout) store_slice(c, c.ys, gout[l], a.ldh, row0, B);
      x = nxt;
      float *t = cur; cur = nxt; nxt = t;
    }
    return x;
  };
  // a = actor(s)
  const float *h = forward(a.actor, sp.xsa, keep_a, a.h_out);
  head_full(h, H, a.actor.Wh, a.actor.ldwh, a.actor.bh, A, true, act);
  // the actor head's backward needs the full last hidden activation later: keep a copy in xns? no --
  // it only needs the LeakyReLU' mask of the own slice (keep_a[L-1]) and the head weights.
  for (int e = tid; e < A * RC; e += CT) sp.xsa[D * RC + e] = act[e];
  __syncthreads();
  // q = critic([s, a]) with the stepped critic
  h = forward(a.c, sp.xsa, keep_c, nullptr);
  head_full(h, H, a.c.Wh, a.c.ldwh, a.c.bh, 1, false, q);
  if (tid == 0 && c.rank == 0) {
    float qs = 0.f;
    for (int r = 0; r < RC; ++r)
      if (row0 + r < B) qs += q[r];
    float *mp = a.metric_partials + size_t(blockIdx.x / CS) * 4;
    mp[0] = -qs; mp[1] = 0.f; mp[2] = qs; mp[3] = 0.f;
  }
  // d(-mean q)/d(critic hidden): last hidden layer (full vector, every CTA)
  const float invB = 1.0f / float(B);
  for (int e = tid; e < H * RC; e += CT) {
    const int k = e / RC, r = e - k * RC;
    const float g = (row0 + r < B) ? -invB * __ldg(a.c.Wh + k) : 0.f;
    nxt[e] = cur[e] > 0.f ? g : g * kLeakySlope;
  }
  c.cluster.sync();      // peers are done reading `cur` before the next layer's remote stores land in it
  { float *t = cur; cur = nxt; nxt = t; }
  for (int l = L - 1; l >= 1; --l, ++s) {
    cluster_layer<EP_DLEAKY>(c, s, cur, nullptr, keep_c[l - 1], nxt);
    float *t = cur; cur = nxt; nxt = t;
  }
  // dq/da through the critic's first layer (action columns), times tanh'  (every CTA, redundantly)
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int p = warp; p < A * RC; p += CT / 32) {
      const int j = p / RC, r = p - j * RC;
      float acc = 0.f;
      for (int n = lane; n < H; n += 32)
        acc = fmaf(cur[n * RC + r], __ldg(a.c.W[0] + size_t(n) * a.c.ldw[0] + D + j), acc);
#pragma unroll
      for (int sh = 16; sh >= 1; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
      if (lane == 0) {
        const float t = act[p];
        const float v = acc * (1.0f - t * t);
        da[p] = v;
        if (c.rank == 0 && row0 + r < B) a.da_out[size_t(row0 + r) * 4 + j] = v;
      }
    }
    if (c.rank == 0)
      for (int p = tid; p < RC; p += CT)
        for (int j = A; j < 4; ++j)
          if (row0 + p < B) a.da_out[size_t(row0 + p) * 4 + j] = 0.f;
  }
  __syncthreads();
  // backward through the actor head: dz of the actor's last hidden layer needs its full activation for the
  // mask; every CTA only has its own slice -> compute the own slice and broadcast it like a layer output
  {
    const int w = cc.w;
    for (int e = tid; e < w * RC; e += CT) {
      const int j = e / RC, r = e - j * RC, k = c.rank * w + j;
      float g = 0.f;
      for (int jj = 0; jj < A; ++jj) g = fmaf(da[jj * RC + r], __ldg(a.actor.Wh + size_t(jj) * a.actor.ldwh + k), g);
      c.ys[e] = keep_a[L - 1][e] > 0.f ? g : g * kLeakySlope;
    }
    __syncthreads();
    const int pieces = (w * RC) >> 2;
    for (int idx = tid; idx < CS * pieces; idx += CT) {
      const int d = idx / pieces, p = idx - d * pieces;
      float *remote = c.cluster.map_shared_rank(nxt, d);
      *reinterpret_cast<float4 *>(remote + c.rank * w * RC + p * 4) = *reinterpret_cast<const float4 *>(c.ys + p * 4);
    }
    store_slice(c, c.ys, a.dz_out[L - 1], a.ldh, row0, B);
    c.cluster.sync();
    float *t = cur; cur = nxt; nxt = t;
  }
  for (int l = L - 1; l >= 1; --l, ++s) {
    cluster_layer<EP_DLEAKY>(c, s, cur, nullptr, keep_a[l - 1], nxt);
    store_slice(c, c.ys, a.dz_out[l - 1], a.ldh, row0, B);
    float *t = cur; cur = nxt; nxt = t;
  }
  cp_async_wait_all();
  c.cluster.sync();
}

int g_smem_set[2] = {0, 0};

template <typename K>
void ensure_smem(K kernel, size_t bytes, int *flag) {
  if (*flag < int(bytes)) {
    GCRL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
    *flag = int(bytes);
  }
}

void fill_common(ClusterCommon &cc, int B, int D, int A, int H, int L, int ldh, int ldc) {
  cc.B = B; cc.D = D; cc.A = A; cc.H = H; cc.L = L; cc.ldh = ldh; cc.ldc = ldc; cc.w = H / CS;
  cc.nsteps = 0;
}
void add_forward(ClusterCommon &cc, const FusedNet &n, int K0) {
  for (int l = 0; l < cc.L; ++l) cc.steps[cc.nsteps++] = MatStep{n.Wt[l], n.ldt[l], l == 0 ? K0 : cc.H};
}
void add_backward(ClusterCommon &cc, const FusedNet &n) {
  for (int l = cc.L - 1; l >= 1; --l) cc.steps[cc.nsteps++] = MatStep{n.W[l], n.ldw[l], cc.H};
}

}  // namespace

bool cluster_supported(int B, int D, int A, int H, int L) {
  if (B < 1 || B > 4096 || L < 1 || L > kFusedMaxL || A > 4 || H < 32 || H > 256 || (H % 32) != 0) return false;
  return cluster_smem_bytes(D, A, H, L) <= size_t(220) * 1024;
}

// one-time host setup outside stream capture
void cluster_init(int D, int A, int H, int L) {
  const size_t smem = cluster_smem_bytes(D, A, H, L);
  ensure_smem(cluster_critic_kernel, smem, &g_smem_set[0]);
  ensure_smem(cluster_actor_kernel, smem, &g_smem_set[1]);
}

int launch_cluster_critic(const FusedCriticArgs &a, cudaStream_t st) {
  ClusterCommon cc;
  fill_common(cc, a.B, a.D, a.A, a.H, a.L, a.ldh, a.ldc);
  add_forward(cc, a.ta, a.D);
  add_forward(cc, a.tc, a.D + a.A);
  add_forward(cc, a.c, a.D + a.A);
  add_backward(cc, a.c);
  const int clusters = (a.B + RC - 1) / RC;
  const size_t smem = cluster_smem_bytes(a.D, a.A, a.H, a.L);
  ensure_smem(cluster_critic_kernel, smem, &g_smem_set[0]);
  cluster_critic_kernel<<<clusters * CS, CT, smem, st>>>(cc, a);
  GCRL_LAUNCHED();
  return clusters;
}

int launch_cluster_actor(const FusedActorArgs &a, cudaStream_t st) {
  ClusterCommon cc;
  fill_common(cc, a.B, a.D, a.A, a.H, a.L, a.ldh, 0);
  add_forward(cc, a.actor, a.D);
  add_forward(cc, a.c, a.D + a.A);
  add_backward(cc, a.c);
  add_backward(cc, a.actor);
  const int clusters = (a.B + RC - 1) / RC;
  const size_t smem = cluster_smem_bytes(a.D, a.A, a.H, a.L);
  ensure_smem(cluster_actor_kernel, smem, &g_smem_set[1]);
  cluster_actor_kernel<<<clusters * CS, CT, smem, st>>>(cc, a);
  GCRL_LAUNCHED();
  return clusters;
}

}  // namespace gcrl
