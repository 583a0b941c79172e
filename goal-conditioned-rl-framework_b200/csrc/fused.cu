// Row-slab fused DDPG update kernels (small / medium batch regime).
//
// Every network pass of the update is independent per batch row; only the weight gradients
// reduce over rows.  So one CTA takes a slab of R batch rows and carries it through ALL layers:
//   critic phase (reference src/agent.py:1302-1329):
//     a' = target_actor(s')  ->  q' = target_critic(s', a')  ->  y = clamp(r + gamma (1-d) q')
//     q = critic(s, a)       ->  dL/dq = 2 (q - y) / B       ->  backward through the critic
//   actor phase (src/agent.py:1288-1294):
//     a = actor(s) -> q = critic(s, a) (stepped critic) -> d(-mean q)/da -> backward through the actor
// with activations in shared memory ([feature][row] layout, so a thread's R row values are one
// 16/32-byte broadcast load) and weights streamed from L2 with coalesced 16-byte loads.  The
// forward uses a transposed copy Wt[in][out] of every weight (maintained by the Adam / Polyak
// kernels), the input-gradient pass uses W[out][in] as stored: in both the reduction index is the
// row index of the matrix, so one routine serves both.  Warps split the reduction range
// (K-split) and combine in shared memory in a fixed order -> deterministic.
// The kernels leave the per-layer activations and pre-activation gradients in global memory; the
// weight-gradient GEMMs of all layers then run as ONE multi-problem launch (mlp.cu).
// Launches per update: 57 -> 9.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "mlp.cuh"

namespace gcrl {

constexpr int kFusedThreads = 512;
constexpr int kFusedWarps = kFusedThreads / 32;

enum : int { EPI_BIAS_LEAKY = 0, EPI_DLEAKY = 1 };

template <int R>
struct RowVec {
  float v[R];
};

template <int R>
__device__ __forceinline__ void load_rows(const float *p, float (&x)[R]) {
  if constexpr (R == 2) {
    const float2 t = *reinterpret_cast<const float2 *>(p);
    x[0] = t.x; x[1] = t.y;
  } else {
    const float4 *q = reinterpret_cast<const float4 *>(p);
#pragma unroll
    for (int i = 0; i < R / 4; ++i) {
      const float4 t = q[i];
      x[4 * i] = t.x; x[4 * i + 1] = t.y; x[4 * i + 2] = t.z; x[4 * i + 3] = t.w;
    }
  }
}

// yT[j][r] = epi( sum_i xT[i][r] * M[i * ldm + j] ),  i in [0, K), j in [0, N)
//   EPI_BIAS_LEAKY: leaky(acc + bias[j])          (forward layer; M = Wt)
//   EPI_DLEAKY    : acc * leaky'(refT[j][r])      (input gradient;  M = W)
// xT, refT, yT: shared memory [feature][R]; red: shared scratch of 2048 * R floats.
// All threads of the CTA must call; ends with a __syncthreads().
template <int R, int EPI>
__device__ __forceinline__ void slab_matmul(const float *xT, int K, const float *__restrict__ M, int ldm, int N,
                                            const float *__restrict__ bias, const float *refT, float *yT,
                                            float *red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ngroups = (N + 3) >> 2;              // 4 adjacent output columns per thread
  const int ncw = (ngroups + 31) >> 5;           // warps needed to cover the columns
  const int ks_total = kFusedWarps / ncw;        // K-split factor
  const int Np = ncw * 128;
  const int cw = warp % ncw, ks = warp / ncw;
  const int j0 = (cw * 32 + lane) * 4;
  if (ks < ks_total) {
    const int kc = (K + ks_total - 1) / ks_total;
    const int kbeg = ks * kc, kend = min(K, kbeg + kc);
    float acc[4][R];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < R; ++r) acc[c][r] = 0.f;
    if (j0 < N) {
      const float *mp = M + size_t(kbeg) * ldm + j0;
#pragma unroll 8
      for (int i = kbeg; i < kend; ++i, mp += ldm) {
        const float4 w = __ldg(reinterpret_cast<const float4 *>(mp));
        float x[R];
        load_rows<R>(xT + i * R, x);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          acc[0][r] = fmaf(w.x, x[r], acc[0][r]);
          acc[1][r] = fmaf(w.y, x[r], acc[1][r]);
          acc[2][r] = fmaf(w.z, x[r], acc[2][r]);
          acc[3][r] = fmaf(w.w, x[r], acc[3][r]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if constexpr (R == 2) {
        *reinterpret_cast<float2 *>(red + (size_t(ks) * Np + j0 + c) * R) = make_float2(acc[c][0], acc[c][1]);
      } else {
        float4 *dst = reinterpret_cast<float4 *>(red + (size_t(ks) * Np + j0 + c) * R);
#pragma unroll
        for (int q = 0; q < R / 4; ++q)
          dst[q] = make_float4(acc[c][4 * q], acc[c][4 * q + 1], acc[c][4 * q + 2], acc[c][4 * q + 3]);
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < N * R; e += kFusedThreads) {
    const int j = e / R, r = e - j * R;
    float s = red[size_t(j) * R + r];
    for (int k2 = 1; k2 < ks_total; ++k2) s += red[(size_t(k2) * Np + j) * R + r];
    if (EPI == EPI_BIAS_LEAKY) {
      s += bias[j];
      s = s > 0.f ? s : s * kLeakySlope;
    } else {
      s = refT[e] > 0.f ? s : s * kLeakySlope;
    }
    yT[e] = s;
  }
  __syncthreads();
}

// out[j][r] = f( sum_k hT[k][r] * Wh[j * ldw + k] + bh[j] ), j < nout <= 4: one warp per (j, r)
template <int R>
__device__ __forceinline__ void slab_head(const float *hT, int K, const float *__restrict__ Wh, int ldw,
                                          const float *__restrict__ bh, int nout, bool tanh_out, float *out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int p = warp; p < nout * R; p += kFusedWarps) {
    const int j = p / R, r = p - j * R;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(hT[k * R + r], __ldg(Wh + size_t(j) * ldw + k), acc);
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) {
      float v = acc + bh[j];
      out[p] = tanh_out ? tanhf(v) : v;
    }
  }
  __syncthreads();
}

// smem [feature][R] -> global [row0 + r][ld] (rows >= B skipped)
template <int R>
__device__ __forceinline__ void store_rows(const float *yT, int N, float *__restrict__ out, int ld, int row0,
                                           int B) {
  for (int e = threadIdx.x; e < N * R; e += kFusedThreads) {
    const int r = e / N, j = e - r * N;
    if (row0 + r < B) out[size_t(row0 + r) * ld + j] = yT[j * R + r];
  }
}

// hidden stack of one network: in (K0 features) -> L layers; layer outputs go to hs[l] when
// keep (all kept) else ping-pong between tmp0/tmp1.  Returns the last layer's buffer.
template <int R>
__device__ __forceinline__ const float *slab_forward(const FusedNet &n, int L, int H, const float *in, int K0,
                                                     float *const *keep, float *tmp0, float *tmp1, float *red,
                                                     float *const *gout, int ldh, int row0, int B) {
  const float *x = in;
  int K = K0;
  for (int l = 0; l < L; ++l) {
    float *y = keep ? keep[l] : ((l & 1) ? tmp1 : tmp0);
    slab_matmul<R, EPI_BIAS_LEAKY>(x, K, n.Wt[l], n.ldt[l], H, n.b[l], nullptr, y, red);
    if (gout) store_rows<R>(y, H, gout[l], ldh, row0, B);
    x = y;
    K = H;
  }
  return x;
}

// Pull a network's parameters into L2 before the layer chain starts.  After an L2 flush (or any other traffic
// that evicted them) every layer step would otherwise begin with a DRAM round trip that all CTAs wait on, since
// they walk the weights in the same order; one 128-byte line per thread, spread over the grid.
__device__ __forceinline__ void prefetch_l2(const float *p, int n, int gtid, int gthreads) {
  for (int e = gtid * 32; e < n; e += gthreads * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + e));
}
__device__ __forceinline__ void prefetch_net(const FusedNet &n, bool forward, bool backward, int gtid, int gthreads) {
  prefetch_l2(n.flat, n.nflat, gtid, gthreads);              // biases, head, W[out][in] (input gradient)
  if (forward) prefetch_l2(n.flatT, n.nflatT, gtid, gthreads);
  (void)backward;
}

struct SmemPlan {
  float *xa, *xb, *t0, *t1, *red, *small;
  float *keep1[kFusedMaxL], *keep2[kFusedMaxL];
};

template <int R>
__device__ __forceinline__ SmemPlan carve(float *base, int KinP, int H, int L, bool two_keeps) {
  SmemPlan p;
  p.xa = base; base += KinP * R;
  p.xb = base; base += KinP * R;
  p.t0 = base; base += H * R;
  p.t1 = base; base += H * R;
  for (int l = 0; l < L; ++l) { p.keep1[l] = base; base += H * R; }
  for (int l = 0; l < L; ++l) { p.keep2[l] = two_keeps ? base : nullptr; if (two_keeps) base += H * R; }
  p.red = base; base += 2048 * R;
  p.small = base;
  return p;
}

size_t fused_smem_bytes(int R, int D, int A, int H, int L, bool two_keeps) {
  const int KinP = (D + A + 3) & ~3;
  size_t f = size_t(2) * KinP * R + size_t(2) * H * R + size_t(L) * H * R * (two_keeps ? 2 : 1) +
             size_t(2048) * R + 24 * R;
  return f * sizeof(float);
}

// ---------------------------------------------------------------------------------------------
// critic phase
// ---------------------------------------------------------------------------------------------
// TD3 = false compiles the smoothing noise / second target critic / given-target / smooth-L1 branches out.
// SPLIT: the slab is carried by a 2-CTA cluster.  The target path (a' = pi_t(s'), q' = Q_t(s', a'): 2L + 2 layer
// steps) and the critic forward (L + 1 steps) do not depend on each other, so CTA 0 runs the former while CTA 1
// runs the latter; CTA 0 drops q' into CTA 1's shared memory (DSMEM), one cluster barrier, and CTA 1 goes on
// with the loss and the backward pass.  Critical path 3L + 3 -> 2L + 2 + (L - 1) layer steps; used when both
// CTAs of every slab fit in one wave.
template <int R, bool TD3, bool SPLIT>
__global__ void __launch_bounds__(kFusedThreads, 1) fused_critic_kernel(FusedCriticArgs a) {
  extern __shared__ float4 fsm4[];
  const int D = a.D, A = a.A, H = a.H, L = a.L, B = a.B;
  const int KinP = (D + A + 3) & ~3;
  SmemPlan sp = carve<R>(reinterpret_cast<float *>(fsm4), KinP, H, L, false);
  pdl_wait();
  pdl_launch_dependents();
  {
    const int gtid = blockIdx.x * kFusedThreads + threadIdx.x, gth = gridDim.x * kFusedThreads;
    if (!TD3 || a.y_in == nullptr) {
      prefetch_net(a.ta, true, false, gtid, gth);
      prefetch_net(a.tc, true, false, gtid, gth);
      if (TD3 && a.has_tc2) prefetch_net(a.tc2, true, false, gtid, gth);
    }
    prefetch_net(a.c, true, true, gtid, gth);
  }
  float *x_ns = sp.xa, *x_sa = sp.xb;
  float *qn = sp.small, *q = qn + R, *yv = q + R, *dzh = yv + R, *rr = dzh + R, *dd = rr + R;
  float *anext = dd + R;                          // [A][R], A <= 4
  const int slab = SPLIT ? int(blockIdx.x >> 1) : int(blockIdx.x);
  const int role = SPLIT ? int(blockIdx.x & 1) : 2;      // 0: target path, 1: critic path, 2: both
  const int row0 = slab * R, tid = threadIdx.x;

  // ---- 0a. (sample) draw this slab's rows from the HER episode store: 3 dependent gathers (bucket record ->
  //          packed row || future offsets -> future goal), relabel + reward in shared memory ----
  const float *tile = nullptr;
  if (a.sample) {
    const HerGeom &g = a.geom;
    float *tl = sp.red;                              // R packed rows (row_f <= 2048 floats each)
    __shared__ uint32_t s_slot[8];
    const SampleScalars sc = *a.sample_sc;
    const int nrows = min(R, B - row0);
    SampleRef ref{0u, 0u, 0u, 0u};
    if (tid < nrows) {
      ref = her_resolve(g, sc, row0 + tid, sc.use_idx ? a.sample_idx : nullptr, nullptr);
      s_slot[tid] = ref.slot;
    }
    __syncthreads();
    const int rf4 = g.row_f >> 2;
    const float4 *rows4 = reinterpret_cast<const float4 *>(g.rows);
    for (int c = tid; c < nrows * rf4; c += kFusedThreads) {
      const int i = c / rf4, q = c - i * rf4;
      reinterpret_cast<float4 *>(tl)[c] = ldg_stream4(rows4 + size_t(s_slot[i]) * rf4 + q);
    }
    float *gfs = sp.small + 16 * R;                  // behind the per-row scalars: [R][8] future goals (G <= 8)
    if (tid < nrows && ref.j > 0) {
      const float *agf = g.ag + size_t(her_future_slot(g, ref)) * g.gpad;
      for (int c = 0; c < g.G; ++c) gfs[tid * 8 + c] = __ldg(agf + c);
    }
    __syncthreads();
    if (tid < nrows && ref.j > 0) her_relabel_row(g, tl + tid * g.row_f, gfs + tid * 8);
    __syncthreads();
    if (role != 0) {                                 // the dense batch, for the actor phase and the second critic
      for (int e = tid; e < nrows * D; e += kFusedThreads) {
        const int i = e / D, k = e - i * D;
        a.bs[size_t(row0 + i) * D + k] = tl[i * g.row_f + k];
        a.bns[size_t(row0 + i) * D + k] = tl[i * g.row_f + g.off_ns + k];
      }
      for (int e = tid; e < nrows * A; e += kFusedThreads) {
        const int i = e / A, k = e - i * A;
        a.ba[size_t(row0 + i) * A + k] = tl[i * g.row_f + g.off_a + k];
      }
      if (tid < nrows) {
        a.br[row0 + tid] = tl[tid * g.row_f + g.off_r];
        a.bd[row0 + tid] = tl[tid * g.row_f + g.off_d];
      }
    }
    tile = tl;
  }
  // ---- 0. stage the slab's rows: x_ns = [s' | (a')], x_sa = [s | a]; also emit [s | a | 0] rows ----
  for (int e = tid; e < KinP * R; e += kFusedThreads) {
    const int k = e / R, r = e - k * R, row = row0 + r;
    float vns = 0.f, vsa = 0.f;
    if (row < B) {
      if (tile != nullptr) {
        const float *t = tile + r * a.geom.row_f;
        if (k < D) { vns = t[a.geom.off_ns + k]; vsa = t[k]; }
        else if (k < D + A) vsa = t[a.geom.off_a + (k - D)];
      } else {
        if (k < D) { vns = a.ns[size_t(row) * D + k]; vsa = a.s[size_t(row) * D + k]; }
        else if (k < D + A) vsa = a.a[size_t(row) * A + (k - D)];
      }
    }
    x_ns[e] = vns;
    x_sa[e] = vsa;
  }
  if (tid < R) {
    const int row = row0 + tid;
    if (tile != nullptr) {
      rr[tid] = row < B ? tile[tid * a.geom.row_f + a.geom.off_r] : 0.f;
      dd[tid] = row < B ? tile[tid * a.geom.row_f + a.geom.off_d] : 0.f;
    } else {
      rr[tid] = row < B ? a.r[row] : 0.f;
      dd[tid] = row < B ? a.d[row] : 0.f;
    }
  }
  __syncthreads();
  if (a.sa_out != nullptr && role != 0) {
    for (int e = tid; e < a.ldc * R; e += kFusedThreads) {
      const int r = e / a.ldc, k = e - r * a.ldc;
      if (row0 + r < B) a.sa_out[size_t(row0 + r) * a.ldc + k] = k < KinP ? x_sa[k * R + r] : 0.f;
    }
  }

  const float *h;
  if ((!TD3 || a.y_in == nullptr) && role != 1) {
    // ---- 1. a' = target_actor(s') (:1312); TD3: + clamp(noise * sigma, +-c), clamped to [-1, 1] (:174-179) ----
    h = slab_forward<R>(a.ta, L, H, x_ns, D, nullptr, sp.t0, sp.t1, sp.red, nullptr, 0, row0, B);
    slab_head<R>(h, H, a.ta.Wh, a.ta.ldwh, a.ta.bh, A, true, anext);
    for (int e = tid; e < A * R; e += kFusedThreads) {
      float v = anext[e];
      if (TD3 && a.noise != nullptr) {
        const int j = e / R, row = row0 + (e - j * R);
        float n = row < B ? a.noise[size_t(row) * A + j] * a.policy_noise : 0.f;
        n = fminf(fmaxf(n, -a.noise_clamp), a.noise_clamp);
        v = fminf(fmaxf(v + n, -1.0f), 1.0f);
      }
      x_ns[D * R + e] = v;
    }
    __syncthreads();
    // ---- 2. q' = target_critic([s', a']) (:1313-1315); TD3: min over the two target critics (:181-183) ----
    h = slab_forward<R>(a.tc, L, H, x_ns, D + A, nullptr, sp.t0, sp.t1, sp.red, nullptr, 0, row0, B);
    slab_head<R>(h, H, a.tc.Wh, a.tc.ldwh, a.tc.bh, 1, false, qn);
    if (TD3 && a.has_tc2) {
      float *qn2 = anext;          // a' already sits in x_ns
      h = slab_forward<R>(a.tc2, L, H, x_ns, D + A, nullptr, sp.t0, sp.t1, sp.red, nullptr, 0, row0, B);
      slab_head<R>(h, H, a.tc2.Wh, a.tc2.ldwh, a.tc2.bh, 1, false, qn2);
      if (tid < R) qn[tid] = fminf(qn[tid], qn2[tid]);
      __syncthreads();
    }
  }
  if (SPLIT) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    if (role == 0) {                     // hand q' to the critic CTA and leave
      float *peer_qn = cluster.map_shared_rank(qn, 1);
      if (tid < R) peer_qn[tid] = qn[tid];
      cluster.sync();
      return;
    }
  }
  // ---- 3. q = critic([s, a]) (:1319), activations kept for the backward pass ----
  h = slab_forward<R>(a.c, L, H, x_sa, D + A, sp.keep1, nullptr, nullptr, sp.red, a.h_out, a.ldh, row0, B);
  slab_head<R>(h, H, a.c.Wh, a.c.ldwh, a.c.bh, 1, false, q);
  if (SPLIT) cooperative_groups::this_cluster().sync();      // q' has arrived
  // ---- 4. Bellman target, loss, dL/dq (:1316-1326) ----
  if (tid == 0) {
    float ls = 0.f, ts = 0.f, qs = 0.f;
    const float invB = 1.0f / float(B);
    for (int r = 0; r < R; ++r) {
      const int row = row0 + r;
      float g = 0.f;
      if (row < B) {
        float y;
        if (TD3 && a.y_in != nullptr) {
          y = a.y_in[row];
        } else {
          y = rr[r] + a.gamma * (1.0f - dd[r]) * qn[r];
          if (a.clamp_y) y = fminf(fmaxf(y, a.y_lo), 0.0f);
        }
        const float diff = q[r] - y;
        const float w = a.is_w != nullptr ? a.is_w[row] : 1.0f;     // (weights * loss).mean(), :1322-1324
        if (!TD3 || a.loss_kind == 0) {               // mse_loss
          ls += w * (diff * diff);
          g = 2.0f * diff * invB * w;
        } else {                                      // smooth_l1_loss, beta = 1 (TD3, :189-197)
          const float ad = fabsf(diff);
          ls += w * (ad < 1.0f ? 0.5f * diff * diff : ad - 0.5f);
          g = (ad < 1.0f ? diff : (diff > 0.f ? 1.0f : -1.0f)) * invB * w;
        }
        float td;
        if (TD3 && a.q_other != nullptr) {            // TD3 critic 2: max of both TD errors, mean of both Q
          const float qo = a.q_other[row];
          td = fmaxf(fabsf(q[r] - y), fabsf(qo - y));
          qs += 0.5f * (q[r] + qo);
        } else {
          td = fabsf(y - q[r]);
          qs += q[r];
        }
        ts += td;
        if (a.td_out != nullptr) a.td_out[row] = td;
        a.dzh_out[row] = g;
        if (a.y_out && (!TD3 || a.y_in == nullptr)) a.y_out[row] = y;
        if (a.q_out) a.q_out[row] = q[r];
      }
      dzh[r] = g;
    }
    float *mp = a.metric_partials + size_t(slab) * 4;
    mp[0] = ls; mp[1] = ts; mp[2] = qs; mp[3] = 0.f;
  }
  __syncthreads();
  // ---- 5. backward through the critic: pre-activation gradients of every hidden layer ----
  float *dz = sp.t0, *dzn = sp.t1;
  for (int e = tid; e < H * R; e += kFusedThreads) {
    const int k = e / R, r = e - k * R;
    const float g = dzh[r] * __ldg(a.c.Wh + k);
    dz[e] = sp.keep1[L - 1][e] > 0.f ? g : g * kLeakySlope;
  }
  __syncthreads();
  store_rows<R>(dz, H, a.dz_out[L - 1], a.ldh, row0, B);
  for (int l = L - 1; l >= 1; --l) {
    slab_matmul<R, EPI_DLEAKY>(dz, H, a.c.W[l], a.c.ldw[l], H, nullptr, sp.keep1[l - 1], dzn, sp.red);
    store_rows<R>(dzn, H, a.dz_out[l - 1], a.ldh, row0, B);
    float *t = dz; dz = dzn; dzn = t;
  }
}

// ---------------------------------------------------------------------------------------------
// actor phase
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kFusedThreads, 1) fused_actor_kernel(FusedActorArgs a) {
  extern __shared__ float4 fsm4[];
  const int D = a.D, A = a.A, H = a.H, L = a.L, B = a.B;
  const int KinP = (D + A + 3) & ~3;
  SmemPlan sp = carve<R>(reinterpret_cast<float *>(fsm4), KinP, H, L, true);
  pdl_wait();
  pdl_launch_dependents();
  {
    const int gtid = blockIdx.x * kFusedThreads + threadIdx.x, gth = gridDim.x * kFusedThreads;
    prefetch_net(a.actor, true, true, gtid, gth);
    prefetch_net(a.c, true, true, gtid, gth);
  }
  float *x_sa = sp.xa;
  float *q = sp.small, *act = q + R, *da = act + 4 * R;   // act, da: [A][R]
  const int row0 = blockIdx.x * R, tid = threadIdx.x;

  for (int e = tid; e < KinP * R; e += kFusedThreads) {
    const int k = e / R, r = e - k * R, row = row0 + r;
    x_sa[e] = (row < B && k < D) ? a.s[size_t(row) * D + k] : 0.f;
  }
  __syncthreads();
  // ---- a = actor(s) (:1289); hidden activations kept (smem) and emitted (global, for wgrad) ----
  const float *h = slab_forward<R>(a.actor, L, H, x_sa, D, sp.keep1, nullptr, nullptr, sp.red, a.h_out, a.ldh,
                                   row0, B);
  slab_head<R>(h, H, a.actor.Wh, a.actor.ldwh, a.actor.bh, A, true, act);
  for (int e = tid; e < A * R; e += kFusedThreads) x_sa[D * R + e] = act[e];
  __syncthreads();
  // ---- q = critic([s, a]) with the stepped critic (:1290) ----
  h = slab_forward<R>(a.c, L, H, x_sa, D + A, sp.keep2, nullptr, nullptr, sp.red, nullptr, 0, row0, B);
  slab_head<R>(h, H, a.c.Wh, a.c.ldwh, a.c.bh, 1, false, q);
  if (tid == 0) {
    float qs = 0.f;
    for (int r = 0; r < R; ++r)
      if (row0 + r < B) qs += q[r];
    float *mp = a.metric_partials + size_t(blockIdx.x) * 4;
    mp[0] = -qs; mp[1] = 0.f; mp[2] = qs; mp[3] = 0.f;     // actor loss = -mean(q) (:1291)
  }
  // ---- d(-mean q) / d(critic hidden), down to the critic's first layer ----
  const float invB = 1.0f / float(B);
  float *dz = sp.t0, *dzn = sp.t1;
  for (int e = tid; e < H * R; e += kFusedThreads) {
    const int k = e / R, r = e - k * R;
    const float g = (row0 + r < B) ? -invB * __ldg(a.c.Wh + k) : 0.f;
    dz[e] = sp.keep2[L - 1][e] > 0.f ? g : g * kLeakySlope;
  }
  __syncthreads();
  for (int l = L - 1; l >= 1; --l) {
    slab_matmul<R, EPI_DLEAKY>(dz, H, a.c.W[l], a.c.ldw[l], H, nullptr, sp.keep2[l - 1], dzn, sp.red);
    float *t = dz; dz = dzn; dzn = t;
  }
  // ---- dq/da through the critic's first layer (action columns), times tanh' ----
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int p = warp; p < A * R; p += kFusedWarps) {
      const int j = p / R, r = p - j * R;
      float acc = 0.f;
      for (int n = lane; n < H; n += 32) acc = fmaf(dz[n * R + r], __ldg(a.c.W[0] + size_t(n) * a.c.ldw[0] + D + j), acc);
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
      if (lane == 0) {
        const float t = act[p];
        const float v = acc * (1.0f - t * t);
        da[p] = v;
        if (row0 + r < B) a.da_out[size_t(row0 + r) * 4 + j] = v;
      }
    }
    for (int p = tid; p < R; p += kFusedThreads)               // zero the unused head columns
      for (int j = A; j < 4; ++j)
        if (row0 + p < B) a.da_out[size_t(row0 + p) * 4 + j] = 0.f;
  }
  __syncthreads();
  // ---- backward through the actor head and hidden stack ----
  for (int e = tid; e < H * R; e += kFusedThreads) {
    const int k = e / R, r = e - k * R;
    float g = 0.f;
    for (int j = 0; j < A; ++j) g = fmaf(da[j * R + r], __ldg(a.actor.Wh + size_t(j) * a.actor.ldwh + k), g);
    dzn[e] = sp.keep1[L - 1][e] > 0.f ? g : g * kLeakySlope;
  }
  __syncthreads();
  { float *t = dz; dz = dzn; dzn = t; }
  store_rows<R>(dz, H, a.dz_out[L - 1], a.ldh, row0, B);
  for (int l = L - 1; l >= 1; --l) {
    slab_matmul<R, EPI_DLEAKY>(dz, H, a.actor.W[l], a.actor.ldw[l], H, nullptr, sp.keep1[l - 1], dzn, sp.red);
    store_rows<R>(dzn, H, a.dz_out[l - 1], a.ldh, row0, B);
    float *t = dz; dz = dzn; dzn = t;
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static int g_fused_smem_set[6][3] = {};

template <typename K>
static void ensure_smem(K kernel, size_t bytes, int *flag) {
  if (*flag < int(bytes)) {
    GCRL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
    *flag = int(bytes);
  }
}

// Rows per CTA.  The kernels are chains of dependent layer steps whose cost grows with R (16 FMAs per weight
// load at R = 4, 8 at R = 2), so the smallest R whose slabs still fit in ONE wave of CTAs wins: at B = 256,
// R = 2 (128 CTAs) runs the critic phase in 36.9 us and the actor phase in ~37 us, R = 4 (64 CTAs) in 48.9 / 45 us
// (36.9 for the critic phase as 2-CTA clusters, see SPLIT).
int fused_rows_per_cta(int B) {
  if (const char *e = getenv("GCRL_FUSED_R")) return atoi(e);
  if (B <= 2 * sm_count()) return 2;
  return B <= 512 ? 4 : 8;
}

// can the critic-phase kernel draw its own rows from this episode store?
bool fused_sample_supported(const HerGeom &g) { return g.G <= 8 && g.row_f <= 2048; }

bool fused_supported(int B, int D, int A, int H, int L) {
  if (B < 1 || B > 1024 || L < 1 || L > kFusedMaxL || A > 4 || H < 4 || H > 2048) return false;
  const int R = fused_rows_per_cta(B);
  if ((B + R - 1) / R > 256) return false;
  return fused_smem_bytes(R, D, A, H, L, true) <= size_t(200) * 1024;
}

static bool fused_split_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("GCRL_FUSED_SPLIT");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int launch_fused_critic(const FusedCriticArgs &a, cudaStream_t st) {
  const int R = fused_rows_per_cta(a.B);
  const int grid = (a.B + R - 1) / R;
  const size_t smem = fused_smem_bytes(R, a.D, a.A, a.H, a.L, false);
  const bool td3 = a.has_tc2 || a.noise != nullptr || a.y_in != nullptr || a.loss_kind != 0 || a.q_other != nullptr;
  // two CTAs per slab (target path | critic path) when there is a target path and everything fits in one wave
  const bool split = fused_split_enabled() && R != 2 && a.y_in == nullptr && 2 * grid <= sm_count();
  auto go = [&](auto kernel, int *flag, bool clustered) {
    ensure_smem(kernel, smem, flag);
    if (!clustered) {
      launch_pdl<PDL_FUSED>(kernel, dim3(grid), dim3(kFusedThreads), smem, st, a);
      return;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * grid);
    cfg.blockDim = dim3(kFusedThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled(PDL_FUSED) ? 2 : 1;
    GCRL_CUDA(cudaLaunchKernelEx(&cfg, kernel, a));
  };
  if (R == 2) {
    if (td3) go(fused_critic_kernel<2, true, false>, &g_fused_smem_set[2][2], false);
    else go(fused_critic_kernel<2, false, false>, &g_fused_smem_set[0][2], false);
  } else if (R == 4) {
    if (td3 && split) go(fused_critic_kernel<4, true, true>, &g_fused_smem_set[5][0], true);
    else if (td3) go(fused_critic_kernel<4, true, false>, &g_fused_smem_set[2][0], false);
    else if (split) go(fused_critic_kernel<4, false, true>, &g_fused_smem_set[4][0], true);
    else go(fused_critic_kernel<4, false, false>, &g_fused_smem_set[0][0], false);
  } else {
    if (td3 && split) go(fused_critic_kernel<8, true, true>, &g_fused_smem_set[5][1], true);
    else if (td3) go(fused_critic_kernel<8, true, false>, &g_fused_smem_set[2][1], false);
    else if (split) go(fused_critic_kernel<8, false, true>, &g_fused_smem_set[4][1], true);
    else go(fused_critic_kernel<8, false, false>, &g_fused_smem_set[0][1], false);
  }
  GCRL_LAUNCHED();
  return grid;
}

int launch_fused_actor(const FusedActorArgs &a, cudaStream_t st) {
  const int R = fused_rows_per_cta(a.B);
  const int grid = (a.B + R - 1) / R;
  const size_t smem = fused_smem_bytes(R, a.D, a.A, a.H, a.L, true);
  if (R == 2) {
    ensure_smem(fused_actor_kernel<2>, smem, &g_fused_smem_set[1][2]);
    launch_pdl<PDL_FUSED>(fused_actor_kernel<2>, dim3(grid), dim3(kFusedThreads), smem, st, a);
  } else if (R == 4) {
    ensure_smem(fused_actor_kernel<4>, smem, &g_fused_smem_set[1][0]);
    launch_pdl<PDL_FUSED>(fused_actor_kernel<4>, dim3(grid), dim3(kFusedThreads), smem, st, a);
  } else {
    ensure_smem(fused_actor_kernel<8>, smem, &g_fused_smem_set[1][1]);
    launch_pdl<PDL_FUSED>(fused_actor_kernel<8>, dim3(grid), dim3(kFusedThreads), smem, st, a);
  }
  GCRL_LAUNCHED();
  return grid;
}

}  // namespace gcrl
