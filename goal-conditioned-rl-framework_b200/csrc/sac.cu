// SAC and "TQC" learners: stochastic BatchNorm actor + truncated critic ensemble.
//
// Replaces (reference src/agent.py) SACAgent.critic_update :548-639, actor_update :513-530,
// alpha_update :532-546, update_critic :487-498, update :659-699 and TQCAgent.critic_update
// :951-1042, actor_update :912-934, alpha_update :936-949, update_critics :888-896, update
// :1062-1100; and SACActorModel.forward / sample (src/model.py:118-141).
//
// One implementation serves both: n critics, the Bellman target and the actor loss use the mean
// of the (n - drop) smallest critic outputs per sample (SAC: n = 2, drop = 1 == torch.min).
// Dense layers run on the fp32 GEMM kernels of mlp.cu; this file adds the BatchNorm1d
// forward / backward (column reductions over the batch), the tanh-Gaussian policy head with
// its log-probability and analytic backward, the sort-truncate-mean over the critic axis (an
// insertion sort in registers), the scalar AdamW on log_alpha and the orchestration.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "mlp.cuh"

struct gcrl_her;
namespace gcrl {
void her_sample_into(gcrl_her *h, int64_t B, const int64_t *idx_host, float *s, float *a, float *r,
                     float *ns, float *d, int64_t *idx_out, cudaStream_t st);
int her_state_dim(const gcrl_her *h);
int her_act_dim(const gcrl_her *h);
int her_device(const gcrl_her *h);
}  // namespace gcrl

using namespace gcrl;

namespace {

inline int pad4(int x) { return (x + 3) & ~3; }
constexpr int kMaxCritics = 8;
constexpr int kSplits = 128;
constexpr int kEnsembleBatchMax = 4096;   // up to this batch the ensemble's backward passes share their launches
constexpr float kBnEps = 1e-5f;        // nn.BatchNorm1d default
constexpr float kBnMomentum = 0.1f;
constexpr float kLogSqrt2Pi = 0.91893853320467274178f;

// ---- BatchNorm1d -----------------------------------------------------------------------------
// One CTA per 32 feature columns, 32 x 32 threads: threadIdx.x = column (coalesced 128-byte row
// segments), threadIdx.y strides the batch rows; fixed-order reduction over threadIdx.y.
__device__ __forceinline__ float col_reduce(float v, float (*red)[33]) {
  red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.y == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += red[i][threadIdx.x];
    red[32][threadIdx.x] = t;
  }
  __syncthreads();
  const float r = red[32][threadIdx.x];
  __syncthreads();
  return r;
}

// z [B][ld] (pre-BN Linear output) is overwritten by xhat; h = relu(xhat * gamma + beta).
__global__ void __launch_bounds__(1024)
bn_fwd_kernel(float *__restrict__ z, int ld, int B, int H, const float *__restrict__ gamma,
              const float *__restrict__ beta, float *__restrict__ rmean, float *__restrict__ rvar,
              float *__restrict__ invstd_out, float *__restrict__ h, int train) {
  __shared__ float red[33][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < H;
  float mu = 0.f, var = 1.f;
  if (train) {
    float s = 0.f;
    if (ok)
      for (int m = threadIdx.y; m < B; m += 32) s += z[size_t(m) * ld + c];
    mu = col_reduce(s, red) / float(B);
    s = 0.f;
    if (ok)
      for (int m = threadIdx.y; m < B; m += 32) {
        const float d = z[size_t(m) * ld + c] - mu;
        s = fmaf(d, d, s);
      }
    var = col_reduce(s, red) / float(B);
    if (ok && threadIdx.y == 0) {   // running statistics: unbiased variance, momentum 0.1
      rmean[c] = (1.0f - kBnMomentum) * rmean[c] + kBnMomentum * mu;
      rvar[c] = (1.0f - kBnMomentum) * rvar[c] + kBnMomentum * (var * (float(B) / float(B - 1)));
    }
  } else if (ok) {
    mu = rmean[c];
    var = rvar[c];
  }
  if (!ok) return;
  const float invstd = 1.0f / sqrtf(var + kBnEps);
  const float g = gamma[c], b = beta[c];
  for (int m = threadIdx.y; m < B; m += 32) {
    const size_t e = size_t(m) * ld + c;
    const float xh = (z[e] - mu) * invstd;
    z[e] = xh;
    h[e] = fmaxf(xh * g + b, 0.f);
  }
  if (threadIdx.y == 0) invstd_out[c] = invstd;
}

// dh [B][ld] = gradient wrt the post-ReLU output, overwritten by the gradient wrt the Linear
// output z.  dgamma / dbeta are written straight into the flat gradient buffer.
__global__ void __launch_bounds__(1024)
bn_bwd_kernel(float *__restrict__ dh, const float *__restrict__ h, const float *__restrict__ xhat, int ld,
              int B, int H, const float *__restrict__ gamma, const float *__restrict__ invstd,
              float *__restrict__ dgamma, float *__restrict__ dbeta) {
  __shared__ float red[33][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < H;
  float s1 = 0.f, s2 = 0.f;
  if (ok)
    for (int m = threadIdx.y; m < B; m += 32) {
      const size_t e = size_t(m) * ld + c;
      const float dy = h[e] > 0.f ? dh[e] : 0.f;
      s1 += dy;
      s2 = fmaf(dy, xhat[e], s2);
    }
  s1 = col_reduce(s1, red);
  s2 = col_reduce(s2, red);
  if (!ok) return;
  const float g = gamma[c], is = invstd[c], n = float(B);
  const float sum_dxhat = s1 * g, sum_dxhat_xhat = s2 * g;
  for (int m = threadIdx.y; m < B; m += 32) {
    const size_t e = size_t(m) * ld + c;
    const float dy = h[e] > 0.f ? dh[e] : 0.f;
    const float dxhat = dy * g;
    dh[e] = is / n * (n * dxhat - sum_dxhat - xhat[e] * sum_dxhat_xhat);
  }
  if (threadIdx.y == 0) {
    dgamma[c] = s2;
    dbeta[c] = s1;
  }
}

// ---- BatchNorm1d with batch statistics over all data-parallel ranks ("sync-BN") --------------------
// The 1-GPU semantics of nn.BatchNorm1d on the concatenated batch (reference src/model.py:103-111):
// every rank writes its local per-column (mean, M2 = sum (z - mean)^2) -- or, in the backward pass, (sum dy,
// sum dy xhat) -- into its slot of a [world][2 ldh] buffer, the caller all-gathers the buffer between two graph
// segments, and every rank merges the slots in rank order (Chan's parallel variance; equal local batch sizes), so
// all ranks normalise with bit-identical statistics.  With world = 1 the result equals bn_fwd_kernel / bn_bwd_kernel
// bit for bit.
__global__ void __launch_bounds__(1024)
bn_stats_kernel(const float *__restrict__ z, int ld, int B, int H, float *__restrict__ slot, int ldh) {
  __shared__ float red[33][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < H;
  float s = 0.f;
  if (ok)
    for (int m = threadIdx.y; m < B; m += 32) s += z[size_t(m) * ld + c];
  const float mu = col_reduce(s, red) / float(B);
  s = 0.f;
  if (ok)
    for (int m = threadIdx.y; m < B; m += 32) {
      const float d = z[size_t(m) * ld + c] - mu;
      s = fmaf(d, d, s);
    }
  const float m2 = col_reduce(s, red);
  if (ok && threadIdx.y == 0) {
    slot[c] = mu;
    slot[ldh + c] = m2;
  }
}

__global__ void __launch_bounds__(1024)
bn_fwd_sync_kernel(float *__restrict__ z, int ld, int B, int H, const float *__restrict__ gamma,
                   const float *__restrict__ beta, float *__restrict__ rmean, float *__restrict__ rvar,
                   float *__restrict__ invstd_out, float *__restrict__ h, const float *__restrict__ gather, int world,
                   int ldh) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  if (c >= H) return;
  float msum = gather[c];
  for (int r = 1; r < world; ++r) msum += gather[size_t(r) * 2 * ldh + c];
  const float mu = msum / float(world);
  float m2 = 0.f;
  for (int r = 0; r < world; ++r) {
    const float d = gather[size_t(r) * 2 * ldh + c] - mu;
    m2 += gather[size_t(r) * 2 * ldh + ldh + c] + float(B) * d * d;
  }
  const float n = float(world) * float(B);
  const float var = m2 / n;
  if (threadIdx.y == 0) {
    rmean[c] = (1.0f - kBnMomentum) * rmean[c] + kBnMomentum * mu;
    rvar[c] = (1.0f - kBnMomentum) * rvar[c] + kBnMomentum * (var * (n / (n - 1.0f)));
  }
  const float invstd = 1.0f / sqrtf(var + kBnEps);
  const float g = gamma[c], b = beta[c];
  for (int m = threadIdx.y; m < B; m += 32) {
    const size_t e = size_t(m) * ld + c;
    const float xh = (z[e] - mu) * invstd;
    z[e] = xh;
    h[e] = fmaxf(xh * g + b, 0.f);
  }
  if (threadIdx.y == 0) invstd_out[c] = invstd;
}

__global__ void __launch_bounds__(1024)
bn_bwd_stats_kernel(const float *__restrict__ dh, const float *__restrict__ h, const float *__restrict__ xhat, int ld,
                    int B, int H, float *__restrict__ slot, int ldh) {
  __shared__ float red[33][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < H;
  float s1 = 0.f, s2 = 0.f;
  if (ok)
    for (int m = threadIdx.y; m < B; m += 32) {
      const size_t e = size_t(m) * ld + c;
      const float dy = h[e] > 0.f ? dh[e] : 0.f;
      s1 += dy;
      s2 = fmaf(dy, xhat[e], s2);
    }
  s1 = col_reduce(s1, red);
  s2 = col_reduce(s2, red);
  if (ok && threadIdx.y == 0) {
    slot[c] = s1;
    slot[ldh + c] = s2;
  }
}

// dgamma / dbeta stay the LOCAL sums (the flat gradient is averaged over the ranks afterwards, like every other
// parameter gradient); the input gradient uses the sums over all ranks and n = world * B rows.
__global__ void __launch_bounds__(1024)
bn_bwd_sync_kernel(float *__restrict__ dh, const float *__restrict__ h, const float *__restrict__ xhat, int ld, int B,
                   int H, const float *__restrict__ gamma, const float *__restrict__ invstd,
                   float *__restrict__ dgamma, float *__restrict__ dbeta, const float *__restrict__ gather, int world,
                   int rank, int ldh) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  if (c >= H) return;
  float s1 = gather[c], s2 = gather[ldh + c];
  for (int r = 1; r < world; ++r) {
    s1 += gather[size_t(r) * 2 * ldh + c];
    s2 += gather[size_t(r) * 2 * ldh + ldh + c];
  }
  const float g = gamma[c], is = invstd[c], n = float(world) * float(B);
  const float sum_dxhat = s1 * g, sum_dxhat_xhat = s2 * g;
  for (int m = threadIdx.y; m < B; m += 32) {
    const size_t e = size_t(m) * ld + c;
    const float dy = h[e] > 0.f ? dh[e] : 0.f;
    const float dxhat = dy * g;
    dh[e] = is / n * (n * dxhat - sum_dxhat - xhat[e] * sum_dxhat_xhat);
  }
  if (threadIdx.y == 0) {
    dgamma[c] = gather[size_t(rank) * 2 * ldh + ldh + c];
    dbeta[c] = gather[size_t(rank) * 2 * ldh + c];
  }
}

// ---- tanh-Gaussian policy head -------------------------------------------------------------------
struct PolicyFwdArgs {
  const float *feat; int ldh;            // last hidden activation [M, K]
  const float *Wm, *bm, *Ws, *bs; int ldw;
  const float *eps;                      // [M, A] or nullptr (deterministic tanh(mean))
  float *rows; int ldr, col0;            // action written into rows[m, col0 + j] (folds torch.cat)
  float *act_out; int ld_act;            // optional dense [M, ld_act] copy (select_action)
  float *logp, *act4, *std4, *gate4;     // [M], [M,4] x 3 (backward cache; may be nullptr)
  int M, K, A;
};

__global__ void __launch_bounds__(256) policy_fwd_kernel(PolicyFwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int m = warp; m < a.M; m += nwarps) {
    const float *h = a.feat + size_t(m) * a.ldh;
    float am[4] = {0.f, 0.f, 0.f, 0.f}, as[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < a.K; k += 32) {
      const float hv = h[k];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < a.A) {
          am[j] = fmaf(hv, a.Wm[size_t(j) * a.ldw + k], am[j]);
          as[j] = fmaf(hv, a.Ws[size_t(j) * a.ldw + k], as[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
        am[j] += __shfl_xor_sync(0xffffffffu, am[j], s);
        as[j] += __shfl_xor_sync(0xffffffffu, as[j], s);
      }
    if (lane == 0) {
      float lp_sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j >= a.A) {
          if (a.act4) { a.act4[size_t(m) * 4 + j] = 0.f; a.std4[size_t(m) * 4 + j] = 0.f; a.gate4[size_t(m) * 4 + j] = 0.f; }
          continue;
        }
        const float mean = am[j] + a.bm[j];
        const float raw = as[j] + a.bs[j];
        const float ls = fminf(fmaxf(raw, -20.0f), 2.0f);       // src/model.py:122
        const float sd = expf(ls);
        float act;
        if (a.eps == nullptr) {
          act = tanhf(mean);
        } else {
          const float xt = mean + sd * a.eps[size_t(m) * a.A + j];   // rsample, :134
          act = tanhf(xt);
          const float diff = xt - mean;
          float lp = -(diff * diff) / (2.0f * (sd * sd)) - logf(sd) - kLogSqrt2Pi;   // Normal.log_prob
          lp -= logf(1.0f - act * act + 1e-8f);                                     // :138
          lp_sum += lp;
        }
        a.rows[size_t(m) * a.ldr + a.col0 + j] = act;
        if (a.act_out) a.act_out[size_t(m) * a.ld_act + j] = act;
        if (a.act4) {
          a.act4[size_t(m) * 4 + j] = act;
          a.std4[size_t(m) * 4 + j] = sd;
          a.gate4[size_t(m) * 4 + j] = (raw >= -20.0f && raw <= 2.0f) ? 1.0f : 0.f;
        }
      }
      if (a.logp) a.logp[m] = lp_sum;
    }
  }
}

// Bellman target: y = r + gamma (1 - d) (mean of the `keep` smallest target-critic values - coef log pi')
__global__ void __launch_bounds__(256)
sac_target_kernel(const float *__restrict__ qt, int64_t ldq, int n, int keep, const float *__restrict__ logp,
                  const float *__restrict__ r, const float *__restrict__ d, float gamma, float coef_const,
                  const float *__restrict__ alpha_dev, float *__restrict__ y, int M) {
  const float coef = coef_const >= 0.f ? coef_const : *alpha_dev;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
    float v[kMaxCritics];
#pragma unroll
    for (int i = 0; i < kMaxCritics; ++i) v[i] = i < n ? qt[int64_t(i) * ldq + m] : 3.4e38f;
#pragma unroll
    for (int i = 1; i < kMaxCritics; ++i) {        // insertion sort, ascending
      const float x = v[i];
      int j = i;
#pragma unroll
      for (int t = 0; t < kMaxCritics; ++t)
        if (j > 0 && v[j - 1] > x) { v[j] = v[j - 1]; --j; }
      v[j] = x;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxCritics; ++i)
      if (i < keep) s += v[i];
    float tq = s / float(keep);
    tq = tq - coef * logp[m];
    y[m] = r[m] + gamma * (1.0f - d[m]) * tq;
  }
}

// Single-CTA fixed-order reductions over the batch (metrics only; B floats per critic).
__device__ __forceinline__ float cta_sum_1024(float v, float *sm) {
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int s = 512; s >= 1; s >>= 1) {
    if (int(threadIdx.x) < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  const float r = sm[0];
  __syncthreads();
  return r;
}

// out[0..n) = mean (q_i - y)^2; out[8] = mean max_i |q_i - y|; out[9] = mean over i, m of q_i
__global__ void __launch_bounds__(1024)
critic_metrics_kernel(const float *__restrict__ q, int64_t ldq, int n, const float *__restrict__ y, int M,
                      float *__restrict__ out, int q_only, const float *__restrict__ is_w, float *__restrict__ td_out) {
  __shared__ float sm[1024];
  float loss[kMaxCritics], td = 0.f, qs = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxCritics; ++i) loss[i] = 0.f;
  for (int m = threadIdx.x; m < M; m += 1024) {
    float mx = 0.f;
    const float yy = q_only ? 0.f : y[m];
    const float w = is_w != nullptr ? is_w[m] : 1.0f;      // prioritised replay: (weights * loss).mean() (:577-581)
#pragma unroll
    for (int i = 0; i < kMaxCritics; ++i)
      if (i < n) {
        const float qq = q[int64_t(i) * ldq + m];
        const float df = qq - yy;
        loss[i] = fmaf(w * df, df, loss[i]);
        mx = fmaxf(mx, fabsf(df));
        qs += qq;
      }
    td += mx;
    if (td_out != nullptr) td_out[m] = mx;                 // max over the critics of |q_i - y| (:620-621, :1021-1024)
  }
  const float inv = 1.0f / float(M);
  if (!q_only) {
#pragma unroll
    for (int i = 0; i < kMaxCritics; ++i)
      if (i < n) {
        const float t = cta_sum_1024(loss[i], sm);
        if (threadIdx.x == 0) out[i] = t * inv;
      }
    const float t = cta_sum_1024(td, sm);
    if (threadIdx.x == 0) out[8] = t * inv;
  }
  const float t = cta_sum_1024(qs, sm);
  if (threadIdx.x == 0) out[9] = t / (float(M) * float(n));
}

// Actor phase: per sample, which critics are among the `keep` smallest -> dq_i = -(1/keep)/M
// (written with stride 4 for head_bwd mode 2) and the actor loss mean(coef log pi - qtrunc).
__global__ void __launch_bounds__(1024)
actor_trunc_kernel(const float *__restrict__ q, int64_t ldq, int n, int keep, const float *__restrict__ logp,
                   float coef_const, const float *__restrict__ alpha_dev, float *__restrict__ dq /*[n][ldq*4]*/,
                   float *__restrict__ loss_out, int M) {
  __shared__ float sm[1024];
  const float coef = coef_const >= 0.f ? coef_const : *alpha_dev;
  const float w = 1.0f / float(keep);
  float ls = 0.f;
  for (int m = threadIdx.x; m < M; m += 1024) {
    float v[kMaxCritics];
    int id[kMaxCritics];
#pragma unroll
    for (int i = 0; i < kMaxCritics; ++i) { v[i] = i < n ? q[int64_t(i) * ldq + m] : 3.4e38f; id[i] = i; }
#pragma unroll
    for (int i = 1; i < kMaxCritics; ++i) {        // stable insertion sort carrying the critic index
      const float x = v[i];
      const int xi = id[i];
      int j = i;
#pragma unroll
      for (int t = 0; t < kMaxCritics; ++t)
        if (j > 0 && v[j - 1] > x) { v[j] = v[j - 1]; id[j] = id[j - 1]; --j; }
      v[j] = x;
      id[j] = xi;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxCritics; ++i)
      if (i < n) {
        const bool kept = i < keep;
        if (kept) s += v[i];
        dq[(int64_t(id[i]) * ldq + m) * 4] = kept ? -(w / float(M)) : 0.f;
      }
    ls += coef * logp[m] - s / float(keep);
  }
  const float t = cta_sum_1024(ls, sm);
  if (threadIdx.x == 0) *loss_out = t / float(M);
}

// dact[m, j] (+)= sum_n dZ0[m, n] W0[n, col0 + j]   (critic layer-1 input gradient, action columns)
__global__ void __launch_bounds__(256)
action_grad_acc_kernel(const float *__restrict__ dZ0, int lddz, const float *__restrict__ W0, int ldw, int col0,
                       float *__restrict__ dact, int M, int N, int nact, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int m = warp; m < M; m += nwarps) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int n = lane; n < N; n += 32) {
      const float g = dZ0[size_t(m) * lddz + n];
      const float *w = W0 + size_t(n) * ldw + col0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nact) acc[j] = fmaf(g, w[j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], s);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = j < nact ? acc[j] : 0.f;
        if (accumulate) v += dact[size_t(m) * 4 + j];
        dact[size_t(m) * 4 + j] = v;
      }
    }
  }
}

// dzh[m][0..4) = dL/dmean, dzh[m][4..8) = dL/d(log_std head output); analytic rsample backward:
// the Normal.log_prob term contributes 0 to dmean and -dlogp to dlog_std.
__global__ void __launch_bounds__(256)
policy_bwd_kernel(const float *__restrict__ dact, const float *__restrict__ act4, const float *__restrict__ std4,
                  const float *__restrict__ gate4, const float *__restrict__ eps, int A, float coef_const,
                  const float *__restrict__ alpha_dev, float *__restrict__ dzh, int M) {
  const float coef = coef_const >= 0.f ? coef_const : *alpha_dev;
  const float dlogp = coef / float(M);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < M * 4; e += gridDim.x * blockDim.x) {
    const int m = e >> 2, j = e & 3;
    float dm = 0.f, dr = 0.f;
    if (j < A) {
      const float a = act4[e];
      const float one_m = 1.0f - a * a;
      const float gx = dact[e] * one_m + dlogp * (2.0f * a * one_m / (one_m + 1e-8f));
      dm = gx;
      dr = (gx * eps[size_t(m) * A + j] * std4[e] - dlogp) * gate4[e];
    }
    dzh[size_t(m) * 8 + j] = dm;
    dzh[size_t(m) * 8 + 4 + j] = dr;
  }
}

// Backward through both policy heads: partial weight / bias gradients per row slab and
// dfeat[m, k] = sum_j dzh[m, j] W[j, k]  (no activation factor: the ReLU mask is applied by bn_bwd).
constexpr int kHeadRows = 256;
__global__ void __launch_bounds__(256)
policy_head_bwd_kernel(const float *__restrict__ dzh, const float *__restrict__ feat, int ldh,
                       const float *__restrict__ Wm, const float *__restrict__ Ws, int ldw, int A,
                       float *__restrict__ dfeat, int lddf, float *__restrict__ pWm, float *__restrict__ pBm,
                       float *__restrict__ pWs, float *__restrict__ pBs, int64_t split_stride, int M, int K,
                       int rows_per_slab) {
  __shared__ float dz_s[kHeadRows][8];
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * rows_per_slab;
  const int nrows = min(rows_per_slab, M - r0);
  for (int i = tid; i < nrows * 8; i += blockDim.x) dz_s[i >> 3][i & 7] = dzh[size_t(r0) * 8 + i];
  __syncthreads();
  if (tid < 8) {
    const int j = tid & 3;
    if (j < A) {
      float s = 0.f;
      for (int i = 0; i < nrows; ++i) s += dz_s[i][tid];
      (tid < 4 ? pBm : pBs)[int64_t(blockIdx.x) * split_stride + j] = s;
    }
  }
  for (int k = tid; k < ldw; k += blockDim.x) {
    float w[8], acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int jj = j & 3;
      w[j] = (k < K && jj < A) ? (j < 4 ? Wm : Ws)[size_t(jj) * ldw + k] : 0.f;
      acc[j] = 0.f;
    }
    if (k < K) {
      for (int i = 0; i < nrows; ++i) {
        const size_t m = size_t(r0 + i);
        const float h = feat[m * ldh + k];
        float dh = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dz = dz_s[i][j];
          dh = fmaf(dz, w[j], dh);
          acc[j] = fmaf(dz, h, acc[j]);
        }
        dfeat[m * lddf + k] = dh;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int jj = j & 3;
      if (jj < A) (j < 4 ? pWm : pWs)[int64_t(blockIdx.x) * split_stride + size_t(jj) * ldw + k] = acc[j];
    }
  }
}

// alpha_update (src/agent.py:532-546): AdamW (wd 0.01, betas 0.9/0.999, eps 1e-8) on the scalar
// log_alpha with gradient -mean(log pi + target_entropy); alpha = exp(log_alpha).  Two kernels so that the
// batch mean can be averaged across data-parallel ranks in between (it sits right behind the actor's
// flat gradient).  st: [0] log_alpha, [1] alpha, [2] m, [3] v
__global__ void __launch_bounds__(1024)
alpha_mean_kernel(const float *__restrict__ logp, int M, float target_entropy, float *__restrict__ mean_out) {
  __shared__ float sm[1024];
  float s = 0.f;
  for (int m = threadIdx.x; m < M; m += 1024) s += logp[m] + target_entropy;
  const float tot = cta_sum_1024(s, sm);
  if (threadIdx.x == 0) *mean_out = tot / float(M);
}

// sc: [0] lr / (1 - b1^t), [1] sqrt(1 - b2^t), [2] 1 - lr * wd  (device scalars, so the update can be a graph)
__global__ void alpha_step_kernel(const float *__restrict__ mean_in, const float *__restrict__ sc,
                                  float *__restrict__ st, float *__restrict__ loss_out) {
  const float step_size = sc[0], bc2_sqrt = sc[1], decay = sc[2];
  const float mean = *mean_in;
  float la = st[0];
  *loss_out = -(la * mean);
  const float g = -mean;
  la *= decay;
  float m1 = st[2], v = st[3];
  m1 = m1 + float(1.0 - 0.9) * (g - m1);
  v = v * float(0.999) + float(1.0 - 0.999) * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + 1e-8f;
  la = la - step_size * (m1 / denom);
  st[0] = la;
  st[1] = expf(la);
  st[2] = m1;
  st[3] = v;
}


// ---- parameter containers ----------------------------------------------------------------------
struct CriticNet {
  int layers = 0;
  std::vector<int> in_d, out_d, ldw, w_off, b_off;
  int total = 0;
  float *p = nullptr, *g = nullptr, *m = nullptr, *v = nullptr;
  void init(int in, int hid, int L, bool trainable, float *grad_store = nullptr) {
    layers = L + 1;
    int off = 0;
    for (int l = 0; l < layers; ++l) {
      in_d.push_back(l == 0 ? in : hid);
      out_d.push_back(l == layers - 1 ? 1 : hid);
      ldw.push_back(pad4(in_d[l]));
      w_off.push_back(off); off += out_d[l] * ldw[l];
      b_off.push_back(off); off += pad4(out_d[l]);
    }
    total = off;
    p = dev_alloc<float>(total);
    GCRL_CUDA(cudaMemset(p, 0, size_t(total) * 4));
    if (trainable) {
      for (float **q : {&m, &v}) { *q = dev_alloc<float>(total); GCRL_CUDA(cudaMemset(*q, 0, size_t(total) * 4)); }
      g = grad_store;          // slice of gcrl_sac::critic_grads (one all-reduce covers the ensemble)
    }
  }
  static int layout_total(int in, int hid, int L) {
    int off = 0;
    for (int l = 0; l <= L; ++l) {
      const int i = l == 0 ? in : hid, o = l == L ? 1 : hid;
      off += o * pad4(i) + pad4(o);
    }
    return off;
  }
  void destroy() { for (float *q : {p, m, v}) if (q) cudaFree(q); }
  const float *W(int l) const { return p + w_off[l]; }
  const float *b(int l) const { return p + b_off[l]; }
};

struct ActorNet {
  int L = 0, D = 0, H = 0, A = 0, ldh = 0;
  std::vector<int> ldw, w_off, b_off, gam_off, bet_off;     // hidden layers
  int wm_off = 0, bm_off = 0, ws_off = 0, bs_off = 0;       // mean / log_std heads [A][ldh]
  int total = 0;
  float *p = nullptr, *g = nullptr, *m = nullptr, *v = nullptr;
  float *rmean = nullptr, *rvar = nullptr;                  // [L][ldh]
  void init(int D_, int H_, int A_, int L_) {
    L = L_; D = D_; H = H_; A = A_; ldh = pad4(H);
    int off = 0;
    for (int l = 0; l < L; ++l) {
      ldw.push_back(pad4(l == 0 ? D : H));
      w_off.push_back(off); off += H * ldw[l];
      b_off.push_back(off); off += ldh;
      gam_off.push_back(off); off += ldh;
      bet_off.push_back(off); off += ldh;
    }
    wm_off = off; off += A * ldh;
    bm_off = off; off += 4;
    ws_off = off; off += A * ldh;
    bs_off = off; off += 4;
    total = off;
    for (float **q : {&p, &g, &m, &v}) {      // g[total] holds mean(log pi + target_entropy) for alpha_update
      *q = dev_alloc<float>(total + 4);
      GCRL_CUDA(cudaMemset(*q, 0, size_t(total + 4) * 4));
    }
    rmean = dev_alloc<float>(size_t(2) * L * ldh);   // running mean | running var, one buffer (data-parallel average)
    rvar = rmean + size_t(L) * ldh;
    std::vector<float> ones(size_t(L) * ldh, 1.0f);
    GCRL_CUDA(cudaMemset(rmean, 0, ones.size() * 4));
    GCRL_CUDA(cudaMemcpy(rvar, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice));
    for (int l = 0; l < L; ++l)   // BatchNorm weight = 1 (torch default)
      GCRL_CUDA(cudaMemcpy(p + gam_off[l], ones.data(), size_t(H) * 4, cudaMemcpyHostToDevice));
  }
  void destroy() { for (float *q : {p, g, m, v, rmean}) if (q) cudaFree(q); }
};

}  // namespace

struct gcrl_sac {
  uint32_t magic = 0x54434153u;   // handle type tag: the two agent families share the Python base class
  int device = 0;
  gcrl_sac_config cfg{};
  int D = 0, A = 0, H = 0, L = 0, n = 0, keep = 0, ldh = 0, ldc = 0;
  int64_t maxB = 0;
  ActorNet actor;
  CriticNet critic[kMaxCritics], target[kMaxCritics];
  int adam_t_c = 0, adam_t_a = 0, adam_t_alpha = 0;
  int dp_B = -1, dp_flags = -1;                  // the update the data-parallel phases belong to
  bool use_graphs = true;
  cudaStream_t cap_stream = nullptr;             // capture-only stream (the caller's may be the legacy one)
  struct GraphRec { cudaGraphExec_t exec; uint64_t kernels; };
  std::map<std::tuple<int, int, int>, GraphRec> graphs;   // (B, flags, phase mask)
  // sync-BN (gcrl_sac_set_sync_bn): the update is cut into graph segments, each ending in the collective the caller
  // runs before the next one (gcrl_sac_update_segment)
  int sync_world = 0, sync_rank = 0;
  float *bn_gather = nullptr;                    // [world][2 ldh]: per-rank BatchNorm partial statistics
  struct SegRec { cudaGraphExec_t exec; uint64_t kernels; int collective; };
  std::map<std::tuple<int, int>, std::vector<SegRec>> seg_graphs;     // (B, flags)
  std::vector<SegRec> *seg_build = nullptr;      // the segment list being captured
  uint64_t seg_mark = 0;
  float *critic_grads = nullptr;                 // [n][critic_stride]: the ensemble's flat gradients, contiguous
  int critic_stride = 0;
  // activations
  std::vector<float *> xhat, ah;                 // actor: [L] x [maxB, ldh]
  float *invstd = nullptr;                       // [L][ldh]
  std::vector<std::vector<float *>> ch;          // critics: [n][L] x [maxB, ldh]
  std::vector<float *> th;                       // target critic scratch [L]
  float *dz[2] = {nullptr, nullptr};
  std::vector<float *> dzc;                      // [2 n] x [min(maxB, kEnsembleBatchMax), ldh]: per-critic gradient ping-pong
  int64_t pset = 0;                              // floats per partial-gradient set (one set per critic)
  float *sa = nullptr, *nsa = nullptr, *spi = nullptr, *br = nullptr, *bd = nullptr;
  float *bs = nullptr, *ba = nullptr, *bns = nullptr, *br0 = nullptr, *bd0 = nullptr;
  float *q = nullptr, *qt = nullptr, *y = nullptr, *dq = nullptr;     // [n][maxB], [n][maxB], [maxB], [n][maxB*4]
  float *logp = nullptr, *act4 = nullptr, *std4 = nullptr, *gate4 = nullptr, *dact = nullptr, *dzh = nullptr;
  float *eps_next = nullptr, *eps_cur = nullptr;
  float *per_w = nullptr, *per_td = nullptr;     // prioritised replay: importance weights in, TD errors out [maxB]
  bool per_on = false;                           // flags bit3 of the update being issued
  float *partials = nullptr;
  int64_t slab = 0;
  float *sumsq = nullptr;
  float *mdev = nullptr;                         // [32] device metrics block
  float *alpha_state = nullptr;                  // [4] log_alpha, alpha, m, v
  StepScalars *d_scalars = nullptr;
  PinnedRing scal_stage, io_stage;
  float *d_io = nullptr;
  size_t io_cap = 0;
};

namespace {

enum : int { M_CLOSS = 0 /*[8]*/, M_TD = 8, M_Q = 9, M_CGN = 10 /*[8]*/, M_ALOSS = 18, M_AGN = 19, M_ALPHA_LOSS = 20 };

int blocks_for(int64_t work, int per_block) {
  return int(std::max<int64_t>(1, std::min<int64_t>((work + per_block - 1) / per_block, int64_t(sm_count()) * 8)));
}

// Collectives a graph segment ends in (gcrl_sac_update_segment): what the caller runs before the next segment.
enum : int { COLL_DONE = 0, COLL_GATHER_BN = 1, COLL_AVG_CRITIC = 2, COLL_AVG_ACTOR = 3 };

// Close the graph segment being captured and open the next one.  Only the segmented capture reaches this.
void seg_cut(gcrl_sac *ag, int collective, bool last = false) {
  cudaGraph_t graph = nullptr;
  GCRL_CUDA(cudaStreamEndCapture(ag->cap_stream, &graph));
  const uint64_t now = launch_counter();
  cudaGraphExec_t exec = nullptr;
  const cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  GCRL_CUDA(e);
  ag->seg_build->push_back(gcrl_sac::SegRec{exec, now - ag->seg_mark, collective});
  ag->seg_mark = now;
  if (!last) GCRL_CUDA(cudaStreamBeginCapture(ag->cap_stream, cudaStreamCaptureModeThreadLocal));
}

// The whole ensemble on the same input rows: one launch per layer (blockIdx.z = critic) instead of n.
// Activations land in ag->ch[i][l] (the caches of the backward pass, or plain scratch for the targets).
void critics_fwd(gcrl_sac *ag, const CriticNet *nets, const float *X, float *q_out, int B, cudaStream_t st) {
  const int n = ag->n, L = ag->L;
  LinearFwdProblem pr[kMaxBatchedLinear];
  for (int l = 0; l < L; ++l) {
    for (int i = 0; i < n; ++i)
      pr[i] = LinearFwdProblem{l == 0 ? X : ag->ch[i][l - 1], l == 0 ? ag->ldc : ag->ldh, nets[i].W(l), nets[i].ldw[l],
                               nets[i].b(l), ag->ch[i][l], ag->ldh};
    launch_linear_fwd_batched(pr, n, B, ag->H, l == 0 ? ag->D + ag->A : ag->H, ACT_LEAKY, st);
  }
  HeadFwdProblem hp[kMaxBatchedLinear];
  for (int i = 0; i < n; ++i)
    hp[i] = HeadFwdProblem{ag->ch[i][L - 1], nets[i].W(L), nets[i].b(L), q_out + int64_t(i) * ag->maxB};
  launch_head_fwd_batched(hp, n, ag->ldh, nets[0].ldw[L], B, ag->H, st);
}

// actor forward on rows[:, :D]; action -> rows[:, D:D+A]
void actor_fwd(gcrl_sac *ag, float *rows, const float *eps, int B, bool train, bool cache, float *act_out,
               cudaStream_t st) {
  ActorNet &a = ag->actor;
  const float *in = rows;
  int ldin = ag->ldc, K = ag->D;
  for (int l = 0; l < ag->L; ++l) {
    launch_linear_fwd(in, ldin, a.p + a.w_off[l], a.ldw[l], a.p + a.b_off[l], ag->xhat[l], ag->ldh, B, ag->H, K,
                      ACT_NONE, st);
    if (train && ag->sync_world > 0) {
      GCRL_REQUIRE(ag->seg_build != nullptr, "a sync-BN agent trains through gcrl_sac_update_segment only");
      bn_stats_kernel<<<(ag->H + 31) / 32, dim3(32, 32), 0, st>>>(
          ag->xhat[l], ag->ldh, B, ag->H, ag->bn_gather + size_t(ag->sync_rank) * 2 * ag->ldh, ag->ldh);
      GCRL_LAUNCHED();
      seg_cut(ag, COLL_GATHER_BN);
      bn_fwd_sync_kernel<<<(ag->H + 31) / 32, dim3(32, 32), 0, st>>>(
          ag->xhat[l], ag->ldh, B, ag->H, a.p + a.gam_off[l], a.p + a.bet_off[l], a.rmean + size_t(l) * ag->ldh,
          a.rvar + size_t(l) * ag->ldh, ag->invstd + size_t(l) * ag->ldh, ag->ah[l], ag->bn_gather, ag->sync_world,
          ag->ldh);
      GCRL_LAUNCHED();
    } else {
      bn_fwd_kernel<<<(ag->H + 31) / 32, dim3(32, 32), 0, st>>>(ag->xhat[l], ag->ldh, B, ag->H, a.p + a.gam_off[l],
                                                                a.p + a.bet_off[l], a.rmean + size_t(l) * ag->ldh,
                                                                a.rvar + size_t(l) * ag->ldh,
                                                                ag->invstd + size_t(l) * ag->ldh, ag->ah[l], train ? 1 : 0);
      GCRL_LAUNCHED();
    }
    in = ag->ah[l]; ldin = ag->ldh; K = ag->H;
  }
  PolicyFwdArgs p{};
  p.feat = ag->ah[ag->L - 1]; p.ldh = ag->ldh;
  p.Wm = a.p + a.wm_off; p.bm = a.p + a.bm_off; p.Ws = a.p + a.ws_off; p.bs = a.p + a.bs_off; p.ldw = ag->ldh;
  p.eps = eps;
  p.rows = rows; p.ldr = ag->ldc; p.col0 = ag->D;
  p.act_out = act_out; p.ld_act = ag->A;
  p.logp = ag->logp;
  if (cache) { p.act4 = ag->act4; p.std4 = ag->std4; p.gate4 = ag->gate4; }
  p.M = B; p.K = ag->H; p.A = ag->A;
  policy_fwd_kernel<<<blocks_for(B, 8), 256, 0, st>>>(p);
  GCRL_LAUNCHED();
}

void reduce_critic(gcrl_sac *ag, CriticNet &c, const int *splits, int head_splits, cudaStream_t st,
                   const float *partials = nullptr) {
  if (partials == nullptr) partials = ag->partials;
  ReduceArgs r{};
  for (int l = 0; l < c.layers; ++l) {
    const int sp = l == c.layers - 1 ? head_splits : splits[l];
    r.seg[r.nseg++] = SegDesc{c.w_off[l], c.out_d[l] * c.ldw[l], partials, sp, ag->slab, c.w_off[l]};
    r.seg[r.nseg++] = SegDesc{c.b_off[l], c.out_d[l], partials, sp, ag->slab, c.b_off[l]};
  }
  r.total = c.total; r.grad = c.g; r.sumsq_partials = ag->sumsq;
  r.metric_partials = nullptr; r.metric_splits = 0; r.metric_scale = 0.f; r.metrics = ag->mdev;
  r.slot_loss = r.slot_td = r.slot_q = -1;
  launch_reduce_grads(r, st);
}

void adam(gcrl_sac *ag, float *p, float *m, float *v, const float *g, int total, int which, float *target,
          bool polyak, int slot_norm, cudaStream_t st) {
  AdamArgs a{};
  a.p = p; a.m = m; a.v = v; a.g = g; a.n = total;
  a.sumsq_partials = ag->sumsq; a.nsumsq = reduce_grid(total);
  a.max_norm = ag->cfg.grad_clip;
  a.weight_decay = ag->cfg.weight_decay;
  a.sc = ag->d_scalars; a.which = which;
  a.target = target; a.tau = ag->cfg.tau; a.one_minus_tau = float(1.0 - double(ag->cfg.tau));
  a.polyak = (polyak && target) ? 1 : 0;
  a.metrics = ag->mdev; a.slot_norm = slot_norm;
  a.tmap = nullptr; a.pT = nullptr; a.targetT = nullptr;
  launch_adam(a, st);
}

// sums of squares of an (averaged) flat gradient that is already in place
void rereduce(gcrl_sac *ag, float *g, int total, cudaStream_t st) {
  ReduceArgs r{};
  r.nseg = 1;
  r.seg[0] = SegDesc{0, total, nullptr, 0, 0, 0};
  r.total = total; r.grad = g; r.sumsq_partials = ag->sumsq;
  r.metrics = ag->mdev; r.slot_loss = r.slot_td = r.slot_q = -1;
  launch_reduce_grads(r, st);
}

enum : int { PH_CGRAD = 1, PH_CSTEP = 2, PH_AGRAD = 4, PH_ASTEP = 8, PH_ALL = 15 };

// critic_update (SAC :548-639, TQC :951-1042).  mask PH_ALL: every critic is stepped right after its own
// backward; data-parallel: PH_CGRAD leaves the n flat gradients in critic_grads (all-reduced by the
// caller), PH_CSTEP clips and steps on the averaged gradients.
void critic_update(gcrl_sac *ag, int B, int flags, int mask, cudaStream_t st) {
  const int n = ag->n, L = ag->L, K0 = ag->D + ag->A;
  const bool tqc = ag->cfg.algo == GCRL_ALGO_TQC;
  const bool dp = mask != PH_ALL;
  if (mask & PH_CGRAD) {
  // next action and log-probability = actor.sample(next_state) in train mode (batch statistics; running statistics move)
  actor_fwd(ag, ag->nsa, ag->eps_next, B, true, false, nullptr, st);
  critics_fwd(ag, ag->target, ag->nsa, ag->qt, B, st);         // ch[i] as scratch: overwritten by the critic forward
  sac_target_kernel<<<blocks_for(B, 256), 256, 0, st>>>(ag->qt, ag->maxB, n, ag->keep, ag->logp, ag->br, ag->bd,
                                                       ag->cfg.gamma, ag->cfg.entropy_coef, ag->alpha_state + 1,
                                                       ag->y, B);
  GCRL_LAUNCHED();
  critics_fwd(ag, ag->critic, ag->sa, ag->q, B, st);
  critic_metrics_kernel<<<1, 1024, 0, st>>>(ag->q, ag->maxB, n, ag->y, B, ag->mdev + M_CLOSS, 0,
                                            ag->per_on ? ag->per_w : nullptr, ag->per_on ? ag->per_td : nullptr);
  GCRL_LAUNCHED();
  if (B <= kEnsembleBatchMax) {
    // the ensemble's backward passes share their launches: one head launch, one weight-gradient and one
    // input-gradient launch per layer (per-critic gradient buffers and partial sets), then reduce + AdamW per critic
    HeadBwdArgs hs[kMaxBatchedLinear];
    for (int i = 0; i < n; ++i) {
      const CriticNet &c = ag->critic[i];
      float *part = ag->partials + int64_t(i) * ag->pset;
      HeadBwdArgs h{};
      h.mode = 0; h.loss_kind = 0; h.clamp_y = 0; h.nout = 1;
      h.q = ag->q + int64_t(i) * ag->maxB; h.y_in = ag->y;
      if (ag->per_on) h.is_w = ag->per_w;
      h.Hact = ag->ch[i][L - 1]; h.ldh = ag->ldh;
      h.W = c.W(L); h.ldw = c.ldw[L];
      h.dZprev = ag->dzc[2 * i]; h.lddz = ag->ldh;
      h.pW = part + c.w_off[L]; h.w_split_stride = ag->slab;
      h.pB = part + c.b_off[L]; h.b_split_stride = ag->slab;
      h.metric_partials = part + ag->slab * kSplits;            // scratch tail, unused
      h.M = B; h.K = ag->H;
      hs[i] = h;
    }
    const int head_splits = launch_head_bwd_batched(hs, n, kSplits, st);
    int splits[8] = {};
    int cur = 0;
    for (int l = L - 1; l >= 0; --l) {
      WgradProblem wp[kMaxWgradProblems];
      for (int i = 0; i < n; ++i) {
        const CriticNet &c = ag->critic[i];
        float *part = ag->partials + int64_t(i) * ag->pset;
        wp[i] = WgradProblem{ag->dzc[2 * i + cur], ag->ldh, l == 0 ? ag->sa : ag->ch[i][l - 1], l == 0 ? ag->ldc : ag->ldh,
                             part + c.w_off[l], c.ldw[l], part + c.b_off[l], ag->H, l == 0 ? K0 : ag->H};
      }
      splits[l] = launch_multi_wgrad(wp, n, B, ag->slab, kSplits, st);
      if (l > 0) {
        LinearDgradProblem dg[kMaxBatchedLinear];
        for (int i = 0; i < n; ++i) {
          const CriticNet &c = ag->critic[i];
          dg[i] = LinearDgradProblem{ag->dzc[2 * i + cur], ag->ldh, c.W(l), c.ldw[l], ag->ch[i][l - 1], ag->ldh,
                                     ag->dzc[2 * i + (cur ^ 1)], ag->ldh};
        }
        launch_linear_dgrad_batched(dg, n, B, ag->H, ag->H, st);
        cur ^= 1;
      }
    }
    for (int i = 0; i < n; ++i) {
      CriticNet &c = ag->critic[i];
      reduce_critic(ag, c, splits, head_splits, st, ag->partials + int64_t(i) * ag->pset);
      if (!dp) adam(ag, c.p, c.m, c.v, c.g, c.total, 0, ag->target[i].p, (flags & 2) != 0, M_CGN + i, st);
    }
  } else
  for (int i = 0; i < n; ++i) {
    CriticNet &c = ag->critic[i];
    HeadBwdArgs h{};
    h.mode = 0; h.loss_kind = 0; h.clamp_y = 0; h.nout = 1;
    h.q = ag->q + int64_t(i) * ag->maxB; h.y_in = ag->y;
    if (ag->per_on) h.is_w = ag->per_w;
    h.Hact = ag->ch[i][L - 1]; h.ldh = ag->ldh;
    h.W = c.W(L); h.ldw = c.ldw[L];
    h.dZprev = ag->dz[0]; h.lddz = ag->ldh;
    h.pW = ag->partials + c.w_off[L]; h.w_split_stride = ag->slab;
    h.pB = ag->partials + c.b_off[L]; h.b_split_stride = ag->slab;
    h.metric_partials = ag->partials + ag->slab * kSplits;      // scratch tail, unused
    h.M = B; h.K = ag->H;
    const int head_splits = launch_head_bwd(h, kSplits, st);
    int splits[8] = {};
    int cur = 0;
    for (int l = L - 1; l >= 0; --l) {
      const float *xin = l == 0 ? ag->sa : ag->ch[i][l - 1];
      const int ldin = l == 0 ? ag->ldc : ag->ldh;
      const int K = l == 0 ? K0 : ag->H;
      splits[l] = launch_linear_wgrad(ag->dz[cur], ag->ldh, xin, ldin, ag->partials + c.w_off[l], c.ldw[l], ag->slab,
                                      ag->partials + c.b_off[l], ag->slab, B, ag->H, K, kSplits, st);
      if (l > 0) {
        launch_linear_dgrad(ag->dz[cur], ag->ldh, c.W(l), c.ldw[l], ag->ch[i][l - 1], ag->ldh, ag->dz[cur ^ 1],
                            ag->ldh, B, ag->H, ag->H, st);
        cur ^= 1;
      }
    }
    reduce_critic(ag, c, splits, head_splits, st);
    if (!dp) adam(ag, c.p, c.m, c.v, c.g, c.total, 0, ag->target[i].p, (flags & 2) != 0, M_CGN + i, st);
  }
  }
  if (ag->seg_build != nullptr) seg_cut(ag, COLL_AVG_CRITIC);
  if (dp && (mask & PH_CSTEP)) {
    for (int i = 0; i < n; ++i) {
      CriticNet &c = ag->critic[i];
      rereduce(ag, c.g, c.total, st);
      adam(ag, c.p, c.m, c.v, c.g, c.total, 0, ag->target[i].p, (flags & 2) != 0, M_CGN + i, st);
    }
  }
  if (tqc && (mask & PH_CSTEP)) {   // logged Q = mean over the STEPPED critics (:1016-1019)
    critics_fwd(ag, ag->critic, ag->sa, ag->qt, B, st);        // the backward pass is done with ch[i]
    critic_metrics_kernel<<<1, 1024, 0, st>>>(ag->qt, ag->maxB, n, nullptr, B, ag->mdev + M_CLOSS, 1, nullptr, nullptr);
    GCRL_LAUNCHED();
  }
}

// actor_update (SAC :513-530, TQC :912-934) + alpha_update (:532-546 / :936-949)
void actor_update(gcrl_sac *ag, int B, int flags, int mask, cudaStream_t st) {
  const int n = ag->n, L = ag->L, D = ag->D, A = ag->A;
  ActorNet &a = ag->actor;
  const bool dp = mask != PH_ALL;
  if (mask & PH_AGRAD) {
  actor_fwd(ag, ag->spi, ag->eps_cur, B, true, true, nullptr, st);
  critics_fwd(ag, ag->critic, ag->spi, ag->q, B, st);
  actor_trunc_kernel<<<1, 1024, 0, st>>>(ag->q, ag->maxB, n, ag->keep, ag->logp, ag->cfg.entropy_coef,
                                        ag->alpha_state + 1, ag->dq, ag->mdev + M_ALOSS, B);
  GCRL_LAUNCHED();
  if (B <= kEnsembleBatchMax) {        // the ensemble's input-gradient chains share their launches
    HeadBwdArgs hs[kMaxBatchedLinear];
    for (int i = 0; i < n; ++i) {
      const CriticNet &c = ag->critic[i];
      HeadBwdArgs h{};
      h.mode = 2; h.nout = 1; h.dz_in = ag->dq + int64_t(i) * ag->maxB * 4;
      h.Hact = ag->ch[i][L - 1]; h.ldh = ag->ldh;
      h.W = c.W(L); h.ldw = c.ldw[L];
      h.dZprev = ag->dzc[2 * i]; h.lddz = ag->ldh;
      h.pW = nullptr; h.pB = nullptr;            // critic weight gradients are discarded
      h.M = B; h.K = ag->H;
      hs[i] = h;
    }
    launch_head_bwd_batched(hs, n, kSplits, st);
    int cur = 0;
    for (int l = L - 1; l >= 1; --l) {
      LinearDgradProblem dg[kMaxBatchedLinear];
      for (int i = 0; i < n; ++i) {
        const CriticNet &c = ag->critic[i];
        dg[i] = LinearDgradProblem{ag->dzc[2 * i + cur], ag->ldh, c.W(l), c.ldw[l], ag->ch[i][l - 1], ag->ldh,
                                   ag->dzc[2 * i + (cur ^ 1)], ag->ldh};
      }
      launch_linear_dgrad_batched(dg, n, B, ag->H, ag->H, st);
      cur ^= 1;
    }
    for (int i = 0; i < n; ++i) {                // summed over the critics in index order
      const CriticNet &c = ag->critic[i];
      action_grad_acc_kernel<<<blocks_for(B, 8), 256, 0, st>>>(ag->dzc[2 * i + cur], ag->ldh, c.W(0), c.ldw[0], D,
                                                              ag->dact, B, ag->H, A, i > 0 ? 1 : 0);
      GCRL_LAUNCHED();
    }
  } else
  for (int i = 0; i < n; ++i) {
    const CriticNet &c = ag->critic[i];
    HeadBwdArgs h{};
    h.mode = 2; h.nout = 1; h.dz_in = ag->dq + int64_t(i) * ag->maxB * 4;
    h.Hact = ag->ch[i][L - 1]; h.ldh = ag->ldh;
    h.W = c.W(L); h.ldw = c.ldw[L];
    h.dZprev = ag->dz[0]; h.lddz = ag->ldh;
    h.pW = nullptr; h.pB = nullptr;              // critic weight gradients are discarded
    h.M = B; h.K = ag->H;
    launch_head_bwd(h, kSplits, st);
    int cur = 0;
    for (int l = L - 1; l >= 1; --l) {
      launch_linear_dgrad(ag->dz[cur], ag->ldh, c.W(l), c.ldw[l], ag->ch[i][l - 1], ag->ldh, ag->dz[cur ^ 1], ag->ldh,
                          B, ag->H, ag->H, st);
      cur ^= 1;
    }
    action_grad_acc_kernel<<<blocks_for(B, 8), 256, 0, st>>>(ag->dz[cur], ag->ldh, c.W(0), c.ldw[0], D, ag->dact, B,
                                                            ag->H, A, i > 0 ? 1 : 0);
    GCRL_LAUNCHED();
  }
  policy_bwd_kernel<<<blocks_for(int64_t(B) * 4, 256), 256, 0, st>>>(ag->dact, ag->act4, ag->std4, ag->gate4,
                                                                    ag->eps_cur, A, ag->cfg.entropy_coef,
                                                                    ag->alpha_state + 1, ag->dzh, B);
  GCRL_LAUNCHED();
  int rows = (B + kSplits - 1) / kSplits;
  rows = std::min(kHeadRows, std::max(rows, 16));
  const int head_slabs = (B + rows - 1) / rows;
  GCRL_REQUIRE(head_slabs <= kSplits, "batch too large for the policy-head partial buffers");
  policy_head_bwd_kernel<<<head_slabs, 256, 0, st>>>(ag->dzh, ag->ah[L - 1], ag->ldh, a.p + a.wm_off, a.p + a.ws_off,
                                                    ag->ldh, A, ag->dz[0], ag->ldh, ag->partials + a.wm_off,
                                                    ag->partials + a.bm_off, ag->partials + a.ws_off,
                                                    ag->partials + a.bs_off, ag->slab, B, ag->H, rows);
  GCRL_LAUNCHED();
  int splits[8] = {};
  int cur = 0;
  for (int l = L - 1; l >= 0; --l) {
    if (ag->sync_world > 0) {
      GCRL_REQUIRE(ag->seg_build != nullptr, "a sync-BN agent trains through gcrl_sac_update_segment only");
      bn_bwd_stats_kernel<<<(ag->H + 31) / 32, dim3(32, 32), 0, st>>>(
          ag->dz[cur], ag->ah[l], ag->xhat[l], ag->ldh, B, ag->H, ag->bn_gather + size_t(ag->sync_rank) * 2 * ag->ldh,
          ag->ldh);
      GCRL_LAUNCHED();
      seg_cut(ag, COLL_GATHER_BN);
      bn_bwd_sync_kernel<<<(ag->H + 31) / 32, dim3(32, 32), 0, st>>>(
          ag->dz[cur], ag->ah[l], ag->xhat[l], ag->ldh, B, ag->H, a.p + a.gam_off[l], ag->invstd + size_t(l) * ag->ldh,
          a.g + a.gam_off[l], a.g + a.bet_off[l], ag->bn_gather, ag->sync_world, ag->sync_rank, ag->ldh);
      GCRL_LAUNCHED();
    } else {
      bn_bwd_kernel<<<(ag->H + 31) / 32, dim3(32, 32), 0, st>>>(ag->dz[cur], ag->ah[l], ag->xhat[l], ag->ldh, B, ag->H,
                                                                a.p + a.gam_off[l], ag->invstd + size_t(l) * ag->ldh,
                                                                a.g + a.gam_off[l], a.g + a.bet_off[l]);
      GCRL_LAUNCHED();
    }
    const float *xin = l == 0 ? ag->spi : ag->ah[l - 1];
    const int ldin = l == 0 ? ag->ldc : ag->ldh;
    const int K = l == 0 ? D : ag->H;
    splits[l] = launch_linear_wgrad(ag->dz[cur], ag->ldh, xin, ldin, ag->partials + a.w_off[l], a.ldw[l], ag->slab,
                                    ag->partials + a.b_off[l], ag->slab, B, ag->H, K, kSplits, st);
    if (l > 0) {
      launch_linear_dgrad(ag->dz[cur], ag->ldh, a.p + a.w_off[l], a.ldw[l], nullptr, 0, ag->dz[cur ^ 1], ag->ldh, B,
                          ag->H, ag->H, st);
      cur ^= 1;
    }
  }
  ReduceArgs r{};
  for (int l = 0; l < L; ++l) {
    r.seg[r.nseg++] = SegDesc{a.w_off[l], ag->H * a.ldw[l], ag->partials, splits[l], ag->slab, a.w_off[l]};
    r.seg[r.nseg++] = SegDesc{a.b_off[l], ag->H, ag->partials, splits[l], ag->slab, a.b_off[l]};
    r.seg[r.nseg++] = SegDesc{a.gam_off[l], ag->H, nullptr, 0, 0, 0};     // written by bn_bwd, already reduced
    r.seg[r.nseg++] = SegDesc{a.bet_off[l], ag->H, nullptr, 0, 0, 0};
  }
  r.seg[r.nseg++] = SegDesc{a.wm_off, A * ag->ldh, ag->partials, head_slabs, ag->slab, a.wm_off};
  r.seg[r.nseg++] = SegDesc{a.bm_off, A, ag->partials, head_slabs, ag->slab, a.bm_off};
  r.seg[r.nseg++] = SegDesc{a.ws_off, A * ag->ldh, ag->partials, head_slabs, ag->slab, a.ws_off};
  r.seg[r.nseg++] = SegDesc{a.bs_off, A, ag->partials, head_slabs, ag->slab, a.bs_off};
  r.total = a.total; r.grad = a.g; r.sumsq_partials = ag->sumsq;
  r.metrics = ag->mdev; r.slot_loss = r.slot_td = r.slot_q = -1;
  launch_reduce_grads(r, st);
  if (flags & 4) {
    alpha_mean_kernel<<<1, 1024, 0, st>>>(ag->logp, B, ag->cfg.target_entropy, a.g + a.total);
    GCRL_LAUNCHED();
  }
  }
  if (ag->seg_build != nullptr) seg_cut(ag, COLL_AVG_ACTOR);
  if (mask & PH_ASTEP) {
    if (dp) rereduce(ag, a.g, a.total, st);
    adam(ag, a.p, a.m, a.v, a.g, a.total, 1, nullptr, false, M_AGN, st);
    if (flags & 4) {
      alpha_step_kernel<<<1, 1, 0, st>>>(a.g + a.total, reinterpret_cast<const float *>(ag->d_scalars + 1),
                                        ag->alpha_state, ag->mdev + M_ALPHA_LOSS);
      GCRL_LAUNCHED();
    } else {
      GCRL_CUDA(cudaMemsetAsync(ag->mdev + M_ALPHA_LOSS, 0, 4, st));
    }
  }
}

void read_metrics(gcrl_sac *ag, int flags, float *metrics_host, cudaStream_t st);

void run_body(gcrl_sac *ag, int B, int flags, int mask, cudaStream_t st) {
  if (mask & (PH_CGRAD | PH_CSTEP)) critic_update(ag, B, flags, mask == PH_ALL ? PH_ALL : (mask & (PH_CGRAD | PH_CSTEP)), st);
  if ((mask & (PH_AGRAD | PH_ASTEP)) && (flags & 1))
    actor_update(ag, B, flags, mask == PH_ALL ? PH_ALL : (mask & (PH_AGRAD | PH_ASTEP)), st);
}

// Replay (or capture on first use) the CUDA graph of (B, flags, phase mask): ~170 launches per TQC update
// become one graph launch; everything that varies per step travels through device scalars.
void run_phases(gcrl_sac *ag, int B, int flags, int mask, cudaStream_t st) {
  // (a graph captured before gcrl_sac_set_sync_bn would otherwise replay with local statistics)
  GCRL_REQUIRE(ag->sync_world == 0, "a sync-BN agent trains through gcrl_sac_update_segment only");
  ag->per_on = (flags & 8) != 0;
  if (!ag->use_graphs) {
    run_body(ag, B, flags, mask, st);
    return;
  }
  const auto key = std::make_tuple(B, flags, mask);
  auto it = ag->graphs.find(key);
  if (it == ag->graphs.end()) {
    cudaGraph_t graph = nullptr;
    const uint64_t before = launch_counter();
    GCRL_CUDA(cudaStreamBeginCapture(ag->cap_stream, cudaStreamCaptureModeThreadLocal));
    try {
      run_body(ag, B, flags, mask, ag->cap_stream);
    } catch (...) {
      cudaStreamEndCapture(ag->cap_stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    GCRL_CUDA(cudaStreamEndCapture(ag->cap_stream, &graph));
    const uint64_t kernels = launch_counter() - before;   // recorded, not executed, by the capture
    count_launch(uint64_t(0) - kernels);
    cudaGraphExec_t exec = nullptr;
    GCRL_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    GCRL_CUDA(cudaGraphDestroy(graph));
    it = ag->graphs.emplace(key, gcrl_sac::GraphRec{exec, kernels}).first;
  }
  GCRL_CUDA(cudaGraphLaunch(it->second.exec, st));
  count_launch(it->second.kernels);
}

// Segmented update (sync-BN): the body is captured ONCE per (B, flags) as a chain of graphs, cut wherever a
// collective has to run (BatchNorm statistics, the two gradient averages); segment `seg` replays graph `seg` and
// reports the collective the caller owes before the next one.
constexpr int PH_SEGMENTED = PH_ALL | 16;          // all four phases, data-parallel form (mask != PH_ALL)

int run_segment(gcrl_sac *ag, int B, int flags, int seg, cudaStream_t st) {
  GCRL_REQUIRE(ag->use_graphs, "sync-BN needs CUDA graphs (unset GCRL_B200_NO_GRAPH)");
  ag->per_on = (flags & 8) != 0;
  const auto key = std::make_tuple(B, flags);
  auto it = ag->seg_graphs.find(key);
  if (it == ag->seg_graphs.end()) {
    std::vector<gcrl_sac::SegRec> segs;
    const uint64_t before = launch_counter();
    ag->seg_build = &segs;
    ag->seg_mark = before;
    GCRL_CUDA(cudaStreamBeginCapture(ag->cap_stream, cudaStreamCaptureModeThreadLocal));
    try {
      run_body(ag, B, flags, PH_SEGMENTED, ag->cap_stream);
      seg_cut(ag, COLL_DONE, true);
    } catch (...) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(ag->cap_stream, &cs);
      if (cs != cudaStreamCaptureStatusNone) {
        cudaGraph_t g = nullptr;
        cudaStreamEndCapture(ag->cap_stream, &g);
        if (g) cudaGraphDestroy(g);
      }
      for (auto &r : segs) cudaGraphExecDestroy(r.exec);
      ag->seg_build = nullptr;
      count_launch(before - launch_counter());
      throw;
    }
    ag->seg_build = nullptr;
    count_launch(before - launch_counter());              // recorded, not executed, by the capture
    it = ag->seg_graphs.emplace(key, std::move(segs)).first;
  }
  GCRL_REQUIRE(seg >= 0 && seg < int(it->second.size()), "segment index past the end of the update");
  const auto &r = it->second[size_t(seg)];
  GCRL_CUDA(cudaGraphLaunch(r.exec, st));
  count_launch(r.kernels);
  return r.collective;
}

// phase < 0: the whole update.  phase 0..3: the data-parallel cut (critic grads | critic steps | actor grads |
// actor step), the caller averaging critic_grads / the actor gradient across ranks in between.
void sac_update(gcrl_sac *ag, int phase, gcrl_her *buf, int64_t B64, const int64_t *idx_host, const float *s,
                const float *a, const float *r, const float *ns, const float *d, const float *eps_next,
                const float *eps_cur, double lr_c, double lr_a, int flags, float *metrics_host, cudaStream_t st,
                int seg = -1, int *collective = nullptr) {
  GCRL_REQUIRE(ag != nullptr, "agent handle is NULL");
  GCRL_REQUIRE(B64 >= 2 && B64 <= ag->maxB, "batch size outside [2, max_batch] (BatchNorm needs > 1 row)");
  const int B = int(B64);
  if (seg > 0) {
    GCRL_REQUIRE(ag->dp_B == B && ag->dp_flags == flags, "segment > 0 must follow segment 0 of the same update");
    *collective = run_segment(ag, B, flags, seg, st);
    return;
  }
  if (phase > 0) {
    GCRL_REQUIRE(ag->dp_B == B && ag->dp_flags == flags, "phase 1..3 must follow phase 0 of the same update");
    run_phases(ag, B, flags, phase == 1 ? PH_CSTEP : (phase == 2 ? PH_AGRAD : PH_ASTEP), st);
    return;
  }
  GCRL_REQUIRE(eps_next != nullptr && (eps_cur != nullptr || !(flags & 1)), "NULL rsample noise tensor");
  if (buf != nullptr) {
    GCRL_REQUIRE(her_state_dim(buf) == ag->D && her_act_dim(buf) == ag->A, "buffer / agent shape mismatch");
    GCRL_REQUIRE(her_device(buf) == ag->device, "buffer and agent live on different devices");
    her_sample_into(buf, B, idx_host, ag->bs, ag->ba, ag->br0, ag->bns, ag->bd0, nullptr, st);
    s = ag->bs; a = ag->ba; r = ag->br0; ns = ag->bns; d = ag->bd0;
  } else {
    GCRL_REQUIRE(s && a && r && ns && d, "NULL batch pointer");
  }
  launch_ingest_batch(s, a, r, ns, d, ag->D, ag->A, B, ag->sa, ag->nsa, ag->spi, ag->ldc, ag->br, ag->bd, st);
  GCRL_CUDA(cudaMemcpyAsync(ag->eps_next, eps_next, size_t(B) * ag->A * 4, cudaMemcpyDeviceToDevice, st));
  if (flags & 1)
    GCRL_CUDA(cudaMemcpyAsync(ag->eps_cur, eps_cur, size_t(B) * ag->A * 4, cudaMemcpyDeviceToDevice, st));
  // per-step optimiser scalars
  auto fill = [&](int t, double lr, float *out) {
    out[0] = float(lr / (1.0 - std::pow(0.9, double(t))));
    out[1] = float(std::sqrt(1.0 - std::pow(0.999, double(t))));
    out[2] = float(1.0 - lr * double(ag->cfg.weight_decay));
    out[3] = 0.f;
  };
  // the step counters move only if the update is issued: a throw below (capture, CUDA error, a misuse check)
  // must not leave the bias corrections one step ahead of the Python schedulers
  const int t_c0 = ag->adam_t_c, t_a0 = ag->adam_t_a, t_al0 = ag->adam_t_alpha;
  try {
  int slot;   // d_scalars[0] = critic / actor AdamW scalars, d_scalars[1] (first 3 floats) = alpha's
  auto *sc = reinterpret_cast<StepScalars *>(ag->scal_stage.acquire(2 * sizeof(StepScalars), &slot));
  ag->adam_t_c += 1;
  fill(ag->adam_t_c, lr_c, &sc->step_size_c);
  if (flags & 1) ag->adam_t_a += 1;
  fill(std::max(1, ag->adam_t_a), lr_a, &sc->step_size_a);
  float *al = reinterpret_cast<float *>(sc + 1);
  al[0] = 0.f; al[1] = 1.f; al[2] = 1.f; al[3] = 0.f;
  if ((flags & 1) && (flags & 4)) {
    ag->adam_t_alpha += 1;
    fill(ag->adam_t_alpha, double(ag->cfg.alpha_lr), al);
  }
  GCRL_CUDA(cudaMemcpyAsync(ag->d_scalars, sc, sizeof(StepScalars) + 16, cudaMemcpyHostToDevice, st));
  ag->scal_stage.release(slot, st);
  if (seg == 0) {
    ag->dp_B = B; ag->dp_flags = flags;
    *collective = run_segment(ag, B, flags, 0, st);
    return;
  }
  if (phase == 0) {
    ag->dp_B = B; ag->dp_flags = flags;
    run_phases(ag, B, flags, PH_CGRAD, st);
    return;
  }
  run_phases(ag, B, flags, PH_ALL, st);
  read_metrics(ag, flags, metrics_host, st);
  } catch (...) {
    ag->adam_t_c = t_c0; ag->adam_t_a = t_a0; ag->adam_t_alpha = t_al0;
    throw;
  }
}

void read_metrics(gcrl_sac *ag, int flags, float *metrics_host, cudaStream_t st) {
  if (metrics_host != nullptr) {
    float hm[32], hs[4];
    GCRL_CUDA(cudaMemcpyAsync(hm, ag->mdev, sizeof(hm), cudaMemcpyDeviceToHost, st));
    GCRL_CUDA(cudaMemcpyAsync(hs, ag->alpha_state, sizeof(hs), cudaMemcpyDeviceToHost, st));
    GCRL_CUDA(cudaStreamSynchronize(st));
    const int n = ag->n;
    if (ag->cfg.algo == GCRL_ALGO_SAC) {
      metrics_host[0] = hm[M_CLOSS]; metrics_host[1] = hm[M_CLOSS + 1];
      metrics_host[5] = hm[M_CGN]; metrics_host[6] = hm[M_CGN + 1];
    } else {                                   // np.mean over the per-critic Python floats (:1013-1014)
      double sl = 0.0, sg = 0.0;
      for (int i = 0; i < n; ++i) { sl += double(hm[M_CLOSS + i]); sg += double(hm[M_CGN + i]); }
      metrics_host[0] = metrics_host[1] = float(sl / n);
      metrics_host[5] = metrics_host[6] = float(sg / n);
    }
    metrics_host[2] = (flags & 1) ? hm[M_ALOSS] : 0.f;
    metrics_host[3] = hm[M_TD];
    metrics_host[4] = hm[M_Q];
    metrics_host[7] = (flags & 1) ? hm[M_AGN] : 0.f;
    metrics_host[8] = (flags & 1) ? hm[M_ALPHA_LOSS] : 0.f;
    metrics_host[9] = hs[1];
    metrics_host[10] = hs[0];
    metrics_host[11] = 0.f;
  }
}

void upload_padded(float *dst, const float *w, int rows, int cols, int ld, cudaStream_t st) {
  std::vector<float> padded(size_t(rows) * ld, 0.f);
  for (int r = 0; r < rows; ++r) std::memcpy(&padded[size_t(r) * ld], w + size_t(r) * cols, size_t(cols) * 4);
  GCRL_CUDA(cudaMemcpyAsync(dst, padded.data(), padded.size() * 4, cudaMemcpyHostToDevice, st));
  GCRL_CUDA(cudaStreamSynchronize(st));
}
void download_padded(float *w, const float *src, int rows, int cols, int ld, cudaStream_t st) {
  std::vector<float> padded(size_t(rows) * ld);
  GCRL_CUDA(cudaMemcpyAsync(padded.data(), src, padded.size() * 4, cudaMemcpyDeviceToHost, st));
  GCRL_CUDA(cudaStreamSynchronize(st));
  for (int r = 0; r < rows; ++r) std::memcpy(w + size_t(r) * cols, &padded[size_t(r) * ld], size_t(cols) * 4);
}

// (rows, cols, ld, weight offset, bias offset) of actor Linear `layer`
void actor_linear_geom(const gcrl_sac *ag, int layer, int *rows, int *cols, int *ld, int *woff, int *boff) {
  const ActorNet &a = ag->actor;
  GCRL_REQUIRE(layer >= 0 && layer <= ag->L + 1, "bad actor layer index");
  if (layer < ag->L) {
    *rows = ag->H; *cols = layer == 0 ? ag->D : ag->H; *ld = a.ldw[layer]; *woff = a.w_off[layer]; *boff = a.b_off[layer];
  } else {
    *rows = ag->A; *cols = ag->H; *ld = ag->ldh;
    *woff = layer == ag->L ? a.wm_off : a.ws_off;
    *boff = layer == ag->L ? a.bm_off : a.bs_off;
  }
}

}  // namespace

// Every entry point checks the handle's type tag: a gcrl_sac handle passed to the other family's functions would be
// reinterpreted as a different struct (garbage shapes and pointers).
static inline void require_handle(const gcrl_sac *h) {
  if (h == nullptr) throw ::gcrl::Error(GCRL_ERR_INVALID, "handle is NULL");
  if (h->magic != 0x54434153u) throw ::gcrl::Error(GCRL_ERR_INVALID, "handle is not a SAC / TQC agent (gcrl_sac_create)");
}

extern "C" {

int gcrl_sac_create(gcrl_sac **out, int device, const gcrl_sac_config *cfg) {
  GCRL_API_BEGIN
  GCRL_REQUIRE(out != nullptr && cfg != nullptr, "NULL argument");
  GCRL_REQUIRE(cfg->algo == GCRL_ALGO_SAC || cfg->algo == GCRL_ALGO_TQC, "unknown algo");
  GCRL_REQUIRE(cfg->state_dim >= 1 && cfg->act_dim >= 1 && cfg->act_dim <= 4, "need state_dim >= 1 and 1 <= act_dim <= 4");
  GCRL_REQUIRE(cfg->hidden_dim >= 1 && cfg->hidden_dim <= 4096, "hidden_dim outside [1, 4096]");
  GCRL_REQUIRE(cfg->layer_count >= 1 && cfg->layer_count <= 6, "layer_count outside [1, 6]");
  GCRL_REQUIRE(cfg->n_critics >= 1 && cfg->n_critics <= kMaxCritics, "n_critics outside [1, 8]");
  GCRL_REQUIRE(cfg->drop_top >= 0 && cfg->drop_top < cfg->n_critics, "drop_top outside [0, n_critics)");
  GCRL_REQUIRE(cfg->max_batch >= 2 && cfg->max_batch <= kSplits * 256, "max_batch outside [2, 32768]");
  GCRL_CUDA(cudaSetDevice(device));
  auto *ag = new gcrl_sac();
  try {
    ag->device = device; ag->cfg = *cfg;
    ag->D = cfg->state_dim; ag->A = cfg->act_dim; ag->H = cfg->hidden_dim; ag->L = cfg->layer_count;
    ag->n = cfg->n_critics; ag->keep = cfg->n_critics - cfg->drop_top;
    ag->ldh = pad4(ag->H); ag->ldc = pad4(ag->D + ag->A); ag->maxB = cfg->max_batch;
    GCRL_REQUIRE(4 * ag->L + 4 <= kMaxSegs, "layer_count too large for the reduction descriptor");
    const int D = ag->D, A = ag->A, H = ag->H, L = ag->L, n = ag->n;
    const size_t mb = size_t(ag->maxB), act = mb * ag->ldh;
    ag->actor.init(D, H, A, L);
    ag->critic_stride = CriticNet::layout_total(D + A, H, L);
    ag->critic_grads = dev_alloc<float>(size_t(n) * ag->critic_stride);
    GCRL_CUDA(cudaMemset(ag->critic_grads, 0, size_t(n) * ag->critic_stride * 4));
    for (int i = 0; i < n; ++i) {
      ag->critic[i].init(D + A, H, L, true, ag->critic_grads + size_t(i) * ag->critic_stride);
      ag->target[i].init(D + A, H, L, false);
    }
    for (int l = 0; l < L; ++l) { ag->xhat.push_back(dev_alloc<float>(act)); ag->ah.push_back(dev_alloc<float>(act)); }
    ag->invstd = dev_alloc<float>(size_t(L) * ag->ldh);
    ag->ch.resize(n);
    for (int i = 0; i < n; ++i)
      for (int l = 0; l < L; ++l) ag->ch[i].push_back(dev_alloc<float>(act));
    for (int l = 0; l < L; ++l) ag->th.push_back(dev_alloc<float>(act));
    for (auto &p : ag->dz) p = dev_alloc<float>(act);
    for (int i = 0; i < 2 * n; ++i) ag->dzc.push_back(dev_alloc<float>(size_t(std::min<int64_t>(mb, kEnsembleBatchMax)) * ag->ldh));
    for (float **p : {&ag->sa, &ag->nsa, &ag->spi}) *p = dev_alloc<float>(mb * ag->ldc);
    for (float **p : {&ag->br, &ag->bd, &ag->br0, &ag->bd0, &ag->y, &ag->logp, &ag->per_w, &ag->per_td}) *p = dev_alloc<float>(mb);
    ag->bs = dev_alloc<float>(mb * D); ag->bns = dev_alloc<float>(mb * D); ag->ba = dev_alloc<float>(mb * A);
    ag->q = dev_alloc<float>(mb * n); ag->qt = dev_alloc<float>(mb * n); ag->dq = dev_alloc<float>(mb * n * 4);
    for (float **p : {&ag->act4, &ag->std4, &ag->gate4, &ag->dact, &ag->eps_next, &ag->eps_cur}) *p = dev_alloc<float>(mb * 4);
    ag->dzh = dev_alloc<float>(mb * 8);
    ag->slab = std::max(ag->actor.total, ag->critic[0].total);
    ag->pset = int64_t(kSplits) * ag->slab + int64_t(kSplits) * 4;
    ag->partials = dev_alloc<float>(size_t(ag->pset) * n);
    ag->sumsq = dev_alloc<float>(size_t(reduce_grid(int(ag->slab))) + 8);
    ag->mdev = dev_alloc<float>(32);
    GCRL_CUDA(cudaMemset(ag->mdev, 0, 32 * 4));
    ag->alpha_state = dev_alloc<float>(4);
    const float init_alpha[4] = {0.f, 1.f, 0.f, 0.f};       // log_alpha = 0 (:424)
    GCRL_CUDA(cudaMemcpy(ag->alpha_state, init_alpha, sizeof(init_alpha), cudaMemcpyHostToDevice));
    ag->d_scalars = dev_alloc<StepScalars>(2);
    GCRL_CUDA(cudaStreamCreateWithFlags(&ag->cap_stream, cudaStreamNonBlocking));
    const char *ng = getenv("GCRL_B200_NO_GRAPH");
    ag->use_graphs = !(ng && ng[0] == '1');
    ag->scal_stage.init(256);
    ag->io_stage.init(size_t(1) << 16);
  } catch (...) {
    delete ag;
    throw;
  }
  *out = ag;
  GCRL_API_END
}

int gcrl_sac_destroy(gcrl_sac *ag) {
  GCRL_API_BEGIN
  if (ag == nullptr) return GCRL_OK;
  cudaSetDevice(ag->device);
  cudaDeviceSynchronize();
  ag->actor.destroy();
  for (int i = 0; i < ag->n; ++i) { ag->critic[i].destroy(); ag->target[i].destroy(); }
  for (auto &v : {ag->xhat, ag->ah, ag->th}) for (float *p : v) cudaFree(p);
  for (auto &v : ag->ch) for (float *p : v) cudaFree(p);
  for (float *p : ag->dzc) cudaFree(p);
  for (float *p : {ag->invstd, ag->dz[0], ag->dz[1], ag->sa, ag->nsa, ag->spi, ag->br, ag->bd, ag->bs, ag->ba, ag->bns,
                   ag->br0, ag->bd0, ag->q, ag->qt, ag->y, ag->dq, ag->logp, ag->act4, ag->std4, ag->gate4, ag->dact,
                   ag->dzh, ag->eps_next, ag->eps_cur, ag->partials, ag->sumsq, ag->mdev, ag->alpha_state, ag->d_io,
                   ag->critic_grads, ag->per_w, ag->per_td})
    if (p) cudaFree(p);
  for (auto &kv : ag->graphs) cudaGraphExecDestroy(kv.second.exec);
  for (auto &kv : ag->seg_graphs) for (auto &r : kv.second) cudaGraphExecDestroy(r.exec);
  if (ag->bn_gather) cudaFree(ag->bn_gather);
  if (ag->cap_stream) cudaStreamDestroy(ag->cap_stream);
  cudaFree(ag->d_scalars);
  ag->scal_stage.destroy();
  ag->io_stage.destroy();
  delete ag;
  GCRL_API_END
}

int gcrl_sac_set_actor_linear(gcrl_sac *ag, int layer, const float *w, const float *b, void *stream) {
  GCRL_API_BEGIN
  require_handle(ag);
  GCRL_REQUIRE(ag && w && b, "NULL argument");
  GCRL_CUDA(cudaSetDevice(ag->device));
  int rows, cols, ld, woff, boff;
  actor_linear_geom(ag, layer, &rows, &cols, &ld, &woff, &boff);
  upload_padded(ag->actor.p + woff, w, rows, cols, ld, as_stream(stream));
  upload_padded(ag->actor.p + boff, b, 1, rows, rows, as_stream(stream));
  GCRL_API_END
}

int gcrl_sac_get_actor_linear(gcrl_sac *ag, int layer, float *w, float *b, void *stream) {
  GCRL_API_BEGIN
  require_handle(ag);
  GCRL_REQUIRE(ag != nullptr, "NULL argument");
  GCRL_CUDA(cudaSetDevice(ag->device));
  int rows, cols, ld, woff, boff;
  actor_linear_geom(ag, layer, &rows, &cols, &ld, &woff, &boff);
  if (w) download_padded(w, ag->actor.p + woff, rows, cols, ld, as_stream(stream));
  if (b) download_padded(b, ag->actor.p + boff, 1, rows, rows, as_stream(stream));
  GCRL_API_END
}

int gcrl_sac_set_actor_bn(gcrl_sac *ag, int layer, const float *weight, const float *bias, const float *rm,
                          const float *rv, void *stream) {
  GCRL_API_BEGIN
  require_handle(ag);
  GCRL_REQUIRE(ag && layer >= 0 && layer < ag->L, "bad BatchNorm layer index");
  GCRL_CUDA(cudaSetDevice(ag->device));
  cudaStream_t st = as_stream(stream);
  ActorNet &a = ag->actor;
  if (weight) upload_padded(a.p + a.gam_off[layer], weight, 1, ag->H, ag->H, st);
  if (bias) upload_padded(a.p + a.bet_off[layer], bias, 1, ag->H, ag->H, st);
  if (rm) upload_padded(a.rmean + size_t(layer) * ag->ldh, rm, 1, ag->H, ag->H, st);
  if (rv) upload_padded(a.rvar + size_t(layer) * ag->ldh, rv, 1, ag->H, ag->H, st);
  GCRL_API_END
}

int gcrl_sac_get_actor_bn(gcrl_sac *ag, int layer, float *weight, float *bias, float *rm, float *rv, void *stream) {
  GCRL_API_BEGIN
  require_handle(ag);
  GCRL_REQUIRE(ag && layer >= 0 && layer < ag->L, "bad BatchNorm layer index");
  GCRL_CUDA(cudaSetDevice(ag->device));
  cudaStream_t st = as_stream(stream);
  ActorNet &a = ag->actor;
  if (weight) download_padded(weight, a.p + a.gam_off[layer], 1, ag->H, ag->H, st);
  if (bias) download_padded(bias, a.p + a.bet_off[layer], 1, ag->H, ag->H, st);
  if (rm) download_padded(rm, a.rmean + size_t(layer) * ag->ldh, 1, ag->H, ag->H, st);
  if (rv) download_padded(rv, a.rvar + size_t(layer) * ag->ldh, 1, ag->H, ag->H, st);
  GCRL_API_END
}

int gcrl_sac_set_critic_layer(gcrl_sac *ag, int critic, int target, int layer, const float *w, const float *b,
                              void *stream) {
  GCRL_API_BEGIN
  require_handle(ag);
  GCRL_REQUIRE(ag && w && b && critic >= 0 && critic < ag->n, "bad critic index / NULL data");
  CriticNet &c = target ? ag->target[critic] : ag->critic[critic];
  GCRL_REQUIRE(layer >= 0 && layer < c.layers, "bad layer index");
  GCRL_CUDA(cudaSetDevice(ag->device));
  upload_padded(c.p + c.w_off[layer], w, c.out_d[layer], c.in_d[layer], c.ldw[layer], as_stream(stream));
  upload_padded(c.p + c.b_off[layer], b, 1, c.out_d[layer], c.out_d[layer], as_stream(stream));
  GCRL_API_END
}

int gcrl_sac_get_critic_layer(gcrl_sac *ag, int critic, int target, int layer, float *w, float *b, void *stream) {
  GCRL_API_BEGIN
  require_handle(ag);
  GCRL_REQUIRE(ag && critic >= 0 && critic < ag->n, "bad critic index");
  CriticNet &c = target ? ag->target[critic] : ag->critic[critic];
  GCRL_REQUIRE(layer >= 0 && layer < c.layers, "bad layer index");
  GCRL_CUDA(cudaSetDevice(ag->device));
  if (w) download_padded(w, c.p + c.w_off[layer], c.out_d[layer], c.in_d[layer], c.ldw[layer], as_stream(stream));
  if (b) download_padded(b, c.p + c.b_off[layer], 1, c.out_d[layer], c.out_d[layer], as_stream(stream));
  GCRL_API_END
}

int gcrl_sac_hard_update(gcrl_sac *ag, void *stream) {
  GCRL_API_BEGIN
  require_handle(ag);
  GCRL_REQUIRE(ag != nullptr, "agent handle is NULL");
  GCRL_CUDA(cudaSetDevice(ag->device));
  for (int i = 0; i < ag->n; ++i)
    GCRL_CUDA(cudaMemcpyAsync(ag->target[i].p, ag->critic[i].p, size_t(ag->critic[i].total) * 4,
                              cudaMemcpyDeviceToDevice, as_stream(stream)));
  GCRL_API_END
}

int gcrl_sac_set_log_alpha(gcrl_sac *ag, float log_alpha, void *stream) {
  GCRL_API_BEGIN
  require_handle(ag);
  GCRL_REQUIRE(ag != nullptr, "agent handle is NULL");
  GCRL_CUDA(cudaSetDevice(ag->device));
  const float v[4] = {log_alpha, std::exp(log_alpha), 0.f, 0.f};   // fresh AdamW state (reset(), :763-765)
  GCRL_CUDA(cudaMemcpyAsync(ag->alpha_state, v, sizeof(v), cudaMemcpyHostToDevice, as_stream(stream)));
  GCRL_CUDA(cudaStreamSynchronize(as_stream(stream)));
  ag->adam_t_alpha = 0;
  GCRL_API_END
}

int gcrl_sac_get_log_alpha(gcrl_sac *ag, float *log_alpha, void *stream) {
  GCRL_API_BEGIN
  require_handle(ag);
  GCRL_REQUIRE(ag != nullptr && log_alpha != nullptr, "NULL argument");
  GCRL_CUDA(cudaSetDevice(ag->device));
  GCRL_CUDA(cudaMemcpyAsync(log_alpha, ag->alpha_state, 4, cudaMemcpyDeviceToHost, as_stream(stream)));
  GCRL_CUDA(cudaStreamSynchronize(as_stream(stream)));
  GCRL_API_END
}

int gcrl_sac_update_batch(gcrl_sac *ag, int64_t B, const float *s, const float *a, const float *r, const float *ns,
                          const float *d, const float *eps_next, const float *eps_cur, double lr_c, double lr_a,
                          int flags, float *metrics_host, void *stream) {
  GCRL_API_BEGIN
  GCRL_NVTX("gcrl_sac_update_batch");
  require_handle(ag);
  GCRL_REQUIRE(ag != nullptr, "agent handle is NULL");
  GCRL_CUDA(cudaSetDevice(ag->device));
  sac_update(ag, -1, nullptr, B, nullptr, s, a, r, ns, d, eps_next, eps_cur, lr_c, lr_a, flags, metrics_host,
             as_stream(stream));
  GCRL_API_END
}

int gcrl_sac_update_from_buffer(gcrl_sac *ag, gcrl_her *buf, int64_t B, const int64_t *idx_host,
                                const float *eps_next, const float *eps_cur, double lr_c, double lr_a, int flags,
                                float *metrics_host, void *stream) {
  GCRL_API_BEGIN
  GCRL_NVTX("gcrl_sac_update_from_buffer");
  require_handle(ag);
  GCRL_REQUIRE(ag != nullptr && buf != nullptr, "NULL handle");
  GCRL_CUDA(cudaSetDevice(ag->device));
  sac_update(ag, -1, buf, B, idx_host, nullptr, nullptr, nullptr, nullptr, nullptr, eps_next, eps_cur, lr_c, lr_a,
             flags, metrics_host, as_stream(stream));
  GCRL_API_END
}

int gcrl_sac_update_phase(gcrl_sac *ag, int phase, gcrl_her *buf, int64_t B, const int64_t *idx_host, const float *s,
                          const float *a, const float *r, const float *ns, const float *d, const float *eps_next,
                          const float *eps_cur, double lr_c, double lr_a, int flags, void *stream) {
  GCRL_API_BEGIN
  GCRL_NVTX("gcrl_sac_update_phase");
  require_handle(ag);
  GCRL_REQUIRE(ag != nullptr && phase >= 0 && phase <= 3, "NULL handle / phase outside 0..3");
  GCRL_CUDA(cudaSetDevice(ag->device));
  sac_update(ag, phase, buf, B, idx_host, s, a, r, ns, d, eps_next, eps_cur, lr_c, lr_a, flags, nullptr,
             as_stream(stream));
  GCRL_API_END
}

int gcrl_sac_set_sync_bn(gcrl_sac *ag, int world, int rank) {
  GCRL_API_BEGIN
  require_handle(ag);
  GCRL_REQUIRE(world >= 0 && world <= 1024 && (world == 0 || (rank >= 0 && rank < world)),
               "need 0 <= world <= 1024 and 0 <= rank < world (world 0 switches sync-BN off)");
  GCRL_CUDA(cudaSetDevice(ag->device));
  GCRL_CUDA(cudaDeviceSynchronize());
  for (auto &kv : ag->seg_graphs) for (auto &r : kv.second) cudaGraphExecDestroy(r.exec);
  ag->seg_graphs.clear();
  if (ag->bn_gather) { cudaFree(ag->bn_gather); ag->bn_gather = nullptr; }
  ag->sync_world = world; ag->sync_rank = world > 0 ? rank : 0;
  if (world > 0) {
    ag->bn_gather = dev_alloc<float>(size_t(world) * 2 * ag->ldh);
    GCRL_CUDA(cudaMemset(ag->bn_gather, 0, size_t(world) * 2 * ag->ldh * 4));
  }
  GCRL_API_END
}

int gcrl_sac_update_segment(gcrl_sac *ag, int segment, gcrl_her *buf, int64_t B, const int64_t *idx_host,
                            const float *s, const float *a, const float *r, const float *ns, const float *d,
                            const float *eps_next, const float *eps_cur, double lr_c, double lr_a, int flags,
                            int *collective, void *stream) {
  GCRL_API_BEGIN
  GCRL_NVTX("gcrl_sac_update_segment");
  require_handle(ag);
  GCRL_REQUIRE(segment >= 0 && collective != nullptr, "segment < 0 / NULL collective");
  GCRL_REQUIRE(ag->sync_world > 0, "gcrl_sac_set_sync_bn first");
  GCRL_CUDA(cudaSetDevice(ag->device));
  sac_update(ag, -1, buf, B, idx_host, s, a, r, ns, d, eps_next, eps_cur, lr_c, lr_a, flags, nullptr,
             as_stream(stream), segment, collective);
  GCRL_API_END
}

int gcrl_sac_per_buffers(gcrl_sac *ag, float **weights_dev, float **td_dev) {
  GCRL_API_BEGIN
  require_handle(ag);
  GCRL_REQUIRE(ag != nullptr && weights_dev != nullptr && td_dev != nullptr, "NULL argument");
  *weights_dev = ag->per_w;
  *td_dev = ag->per_td;
  GCRL_API_END
}

int gcrl_sac_dp_buffer(gcrl_sac *ag, int which, float **dev, int64_t *count) {
  GCRL_API_BEGIN
  require_handle(ag);
  GCRL_REQUIRE(ag != nullptr && dev != nullptr && count != nullptr, "NULL argument");
  switch (which) {
    case 0: *dev = ag->actor.g; *count = ag->actor.total + 4; break;                       // + alpha's batch mean
    case 1: *dev = ag->critic_grads; *count = int64_t(ag->n) * ag->critic_stride; break;
    case 2: *dev = ag->actor.rmean; *count = int64_t(2) * ag->L * ag->ldh; break;          // BatchNorm running stats
    case 3: *dev = ag->mdev; *count = 32; break;
    case 4:
      GCRL_REQUIRE(ag->bn_gather != nullptr, "gcrl_sac_set_sync_bn first");
      *dev = ag->bn_gather; *count = int64_t(ag->sync_world) * 2 * ag->ldh; break;          // [world][2 ldh]
    default:
      throw Error(GCRL_ERR_INVALID,
                  "which must be 0 (actor grad), 1 (critic grads), 2 (BN running stats), 3 (metrics), 4 (sync-BN slots)");
  }
  GCRL_API_END
}

int gcrl_sac_read_metrics(gcrl_sac *ag, int flags, float *metrics_host, void *stream) {
  GCRL_API_BEGIN
  require_handle(ag);
  GCRL_REQUIRE(ag != nullptr && metrics_host != nullptr, "NULL argument");
  GCRL_CUDA(cudaSetDevice(ag->device));
  read_metrics(ag, flags, metrics_host, as_stream(stream));
  GCRL_API_END
}

int gcrl_sac_act(gcrl_sac *ag, int64_t n, const float *obs_host, const float *eps_host, float *act_host,
                 void *stream) {
  GCRL_API_BEGIN
  require_handle(ag);
  GCRL_REQUIRE(ag && obs_host && act_host, "NULL argument");
  GCRL_REQUIRE(n >= 1 && n <= ag->maxB, "row count outside [1, max_batch]");
  GCRL_CUDA(cudaSetDevice(ag->device));
  cudaStream_t st = as_stream(stream);
  const int D = ag->D, A = ag->A;
  const size_t need = size_t(n) * (D + 2 * A);
  if (need > ag->io_cap) {
    GCRL_CUDA(cudaStreamSynchronize(st));
    if (ag->d_io) GCRL_CUDA(cudaFree(ag->d_io));
    ag->io_cap = need * 2;
    ag->d_io = dev_alloc<float>(ag->io_cap);
  }
  auto stage = [&](const float *host, size_t count, size_t off) {
    int slot;
    char *p = ag->io_stage.acquire(count * 4, &slot);
    std::memcpy(p, host, count * 4);
    GCRL_CUDA(cudaMemcpyAsync(ag->d_io + off, p, count * 4, cudaMemcpyHostToDevice, st));
    ag->io_stage.release(slot, st);
    return ag->d_io + off;
  };
  float *obs = stage(obs_host, size_t(n) * D, 0);
  float *eps = eps_host ? stage(eps_host, size_t(n) * A, size_t(n) * D) : nullptr;
  float *out = ag->d_io + size_t(n) * (D + A);
  launch_pack_rows(obs, D, nullptr, A, ag->spi, ag->ldc, int(n), st);
  actor_fwd(ag, ag->spi, eps, int(n), false, false, out, st);     // set_eval(): running statistics
  GCRL_CUDA(cudaMemcpyAsync(act_host, out, size_t(n) * A * 4, cudaMemcpyDeviceToHost, st));
  GCRL_CUDA(cudaStreamSynchronize(st));
  GCRL_API_END
}

}  // extern "C"
