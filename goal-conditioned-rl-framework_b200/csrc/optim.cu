// Gradient reduction, global-norm clipping, Adam / AdamW and Polyak kernels.
//
// Replaces (reference src/agent.py): clip_grad_norm_ (:1295,:1331), get_gradient_norm
// (:1279-1286, 8 host syncs per call -> one device scalar), torch.optim.Adam.step (:1297,
// :1333; AdamW for TD3 :46-48) and the per-parameter Polyak loop (:1259-1271).  Each network
// is one flat fp32 buffer, so an optimiser step is two launches: a fixed-order reduction of
// the split-batch partial gradients (+ sum of squares) and one fused clip + Adam (+ Polyak)
// pass.  Bound: HBM traffic of 7 fp32 words per parameter (g, m, v, p read; m, v, p write).
#include <algorithm>

#include "mlp.cuh"

namespace gcrl {

constexpr int kOptThreads = 256;

int reduce_grid(int total) {
  return std::max(1, std::min((total + kOptThreads - 1) / kOptThreads, sm_count() * 4));
}

__device__ __forceinline__ float block_sum_fixed(float v, float *scratch /*[32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? scratch[lane] : 0.f;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) r += __shfl_xor_sync(0xffffffffu, r, s);
  }
  return r;  // valid in warp 0
}

__device__ __forceinline__ void reduce_grads_body(const ReduceArgs &a) {
  __shared__ float scratch[32];
  float sq = 0.f;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < a.total; e += gridDim.x * blockDim.x) {
    float g = 0.f;
    bool found = false;
#pragma unroll 4
    for (int s = 0; s < a.nseg; ++s) {
      const SegDesc &sd = a.seg[s];
      if (e >= sd.begin && e < sd.begin + sd.count) {
        found = true;
        if (sd.partials != nullptr) {
          const float *src = sd.partials + sd.offset + (e - sd.begin);
          for (int k = 0; k < sd.splits; ++k) g += src[int64_t(k) * sd.split_stride];
        } else {
          g = a.grad[e];
        }
      }
    }
    if (!found) g = 0.f;  // alignment padding between segments
    a.grad[e] = g;
    sq = fmaf(g, g, sq);
  }
  const float tot = block_sum_fixed(sq, scratch);
  if (threadIdx.x == 0) a.sumsq_partials[blockIdx.x] = tot;
  if (blockIdx.x == 0 && a.metric_partials != nullptr && threadIdx.x < 3) {
    float s = 0.f;
    for (int k = 0; k < a.metric_splits; ++k) s += a.metric_partials[size_t(k) * 4 + threadIdx.x];
    const int slot = threadIdx.x == 0 ? a.slot_loss : (threadIdx.x == 1 ? a.slot_td : a.slot_q);
    if (slot >= 0) a.metrics[slot] = s * a.metric_scale;
  }
}

__global__ void __launch_bounds__(kOptThreads) reduce_grads_kernel(ReduceArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  reduce_grads_body(a);
}

void launch_reduce_grads(const ReduceArgs &a, cudaStream_t st) {
  launch_pdl<PDL_OPTIM>(reduce_grads_kernel, dim3(reduce_grid(a.total)), dim3(kOptThreads), 0, st, a);
  GCRL_LAUNCHED();
}

// ---- data-parallel gradient averaging over NVLink peer memory (one process per GPU) ----------------
// Every rank maps the flat gradient buffers of all its peers (CUDA IPC) and averages them itself, in rank
// order 0..world-1 on every rank -> bit-identical replicas, no NCCL launch on the critical path (a 0.55 MB
// all-reduce is pure latency: 32 us through NCCL at 8 GPUs, measured).  Cross-GPU ordering is a flag barrier:
// rank r bumps slot r of every peer's flag array, then spins on its own array.
__global__ void __launch_bounds__(32)
p2p_barrier_kernel(unsigned int *const *peer_flags, unsigned int *epoch, int rank, int world, int *err) {
  __shared__ unsigned int e_s;
  if (threadIdx.x == 0) e_s = *epoch + 1u;
  __syncwarp();
  const unsigned int e = e_s;
  if (int(threadIdx.x) < world) {
    __threadfence_system();                                   // my gradients are visible before the flag is
    volatile unsigned int *dst = peer_flags[threadIdx.x] + rank;
    *dst = e;
    volatile unsigned int *src = peer_flags[rank] + threadIdx.x;
    const long long t0 = clock64();
    while (*src < e) {
      if (clock64() - t0 > 120000000000ll) {                  // ~60 s: a peer died; fail loudly instead of hanging
        *err = 1;
        break;
      }
    }
    __threadfence_system();
  }
  __syncwarp();
  if (threadIdx.x == 0) *epoch = e;
}

void launch_p2p_barrier(unsigned int *const *peer_flags, unsigned int *epoch, int rank, int world, int *err,
                        cudaStream_t st) {
  p2p_barrier_kernel<<<1, 32, 0, st>>>(peer_flags, epoch, rank, world, err);
  GCRL_LAUNCHED();
}

// out[e] = (sum_r peers[r][e]) / world  (+ per-CTA sums of squares for the global-norm clip)
__global__ void __launch_bounds__(kOptThreads)
p2p_reduce_kernel(const float *const *peers, int world, float inv_world, float *__restrict__ out, int n,
                  float *__restrict__ sumsq_partials) {
  __shared__ float scratch[32];
  float sq = 0.f;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) s += __ldcv(peers[r] + e);        // never from a stale L1 line
    const float g = s * inv_world;
    out[e] = g;
    sq = fmaf(g, g, sq);
  }
  const float tot = block_sum_fixed(sq, scratch);
  if (threadIdx.x == 0) sumsq_partials[blockIdx.x] = tot;
}

void launch_p2p_reduce(const float *const *peers, int world, float *out, int n, float *sumsq_partials,
                       cudaStream_t st) {
  p2p_reduce_kernel<<<reduce_grid(n), kOptThreads, 0, st>>>(peers, world, 1.0f / float(world), out, n, sumsq_partials);
  GCRL_LAUNCHED();
}

// metrics in ONE single-block launch: publish the local 8 floats, flag barrier, average the ranks' copies
// (losses, td error and q are batch means; the gradient norms are already global)
__global__ void __launch_bounds__(32)
p2p_metrics_kernel(unsigned int *const *peer_flags, unsigned int *epoch, int rank, int world, int *err,
                   const float *__restrict__ local, float *__restrict__ outbox, const float *const *peer_outbox,
                   float *__restrict__ avg) {
  __shared__ unsigned int e_s;
  if (threadIdx.x < 8) outbox[threadIdx.x] = local[threadIdx.x];
  if (threadIdx.x == 0) e_s = *epoch + 1u;
  __syncwarp();
  const unsigned int e = e_s;
  if (int(threadIdx.x) < world) {
    __threadfence_system();
    volatile unsigned int *dst = peer_flags[threadIdx.x] + rank;
    *dst = e;
    volatile unsigned int *src = peer_flags[rank] + threadIdx.x;
    const long long t0 = clock64();
    while (*src < e) {
      if (clock64() - t0 > 120000000000ll) {
        *err = 1;
        break;
      }
    }
    __threadfence_system();
  }
  __syncwarp();
  if (threadIdx.x < 8) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) s += __ldcv(peer_outbox[r] + threadIdx.x);
    avg[threadIdx.x] = s / float(world);
  }
  if (threadIdx.x == 0) *epoch = e;
}
void launch_p2p_metrics(unsigned int *const *peer_flags, unsigned int *epoch, int rank, int world, int *err,
                        const float *local, float *outbox, const float *const *peer_outbox, float *avg,
                        cudaStream_t st) {
  p2p_metrics_kernel<<<1, 32, 0, st>>>(peer_flags, epoch, rank, world, err, local, outbox, peer_outbox, avg);
  GCRL_LAUNCHED();
}

// torch.optim.Adam(W) single-tensor semantics (betas 0.9/0.999, eps 1e-8):
//   p *= 1 - lr*wd (AdamW only);  m += (1-b1)(g-m);  v = v*b2 + (1-b2) g g;
//   p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// preceded by clip_grad_norm_: g *= min(1, max_norm / (||g|| + 1e-6)).
__device__ __forceinline__ void adam_body(const AdamArgs &a) {
  __shared__ float s_coef, s_norm;
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x < 32) {
    float s = 0.f;
    for (int i = threadIdx.x; i < a.nsumsq; i += 32) s += a.sumsq_partials[i];
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    if (threadIdx.x == 0) {
      const float norm = sqrtf(s);
      float coef = 1.0f;
      if (a.max_norm >= 0.f) coef = fminf(a.max_norm / (norm + 1e-6f), 1.0f);
      s_coef = coef;
      s_norm = norm;
    }
  }
  __syncthreads();
  const float coef = s_coef;
  const float step_size = a.which == 0 ? a.sc->step_size_c : a.sc->step_size_a;
  const float bc2_sqrt = a.which == 0 ? a.sc->bc2_sqrt_c : a.sc->bc2_sqrt_a;
  const float decay = a.which == 0 ? a.sc->decay_c : a.sc->decay_a;
  // scalar arguments exactly as torch narrows them: float(1 - 0.9), float(0.999), float(1 - 0.999)
  const float omb1 = float(1.0 - 0.9), b2 = float(0.999), omb2 = float(1.0 - 0.999), eps = 1e-8f;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < a.n; e += gridDim.x * blockDim.x) {
    const float g = a.g[e] * coef;
    float p = a.p[e];
    if (a.weight_decay != 0.f) p *= decay;
    float m = a.m[e], v = a.v[e];
    m = m + omb1 * (g - m);
    v = v * b2;
    v = v + omb2 * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - step_size * (m / denom);
    a.m[e] = m;
    a.v[e] = v;
    a.p[e] = p;
    const int tm = a.tmap != nullptr ? a.tmap[e] : -1;
    if (tm >= 0) a.pT[tm] = p;
    if (a.polyak) {
      const float t = a.tau * p + a.one_minus_tau * a.target[e];
      a.target[e] = t;
      if (tm >= 0) a.targetT[tm] = t;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.metrics != nullptr && a.slot_norm >= 0)
    a.metrics[a.slot_norm] = s_norm * coef;  // norm after clipping (src/agent.py:1332,:1300)
}

__global__ void __launch_bounds__(kOptThreads) adam_kernel(AdamArgs a) { adam_body(a); }

void launch_adam(const AdamArgs &a, cudaStream_t st) {
  launch_pdl<PDL_OPTIM>(adam_kernel, dim3(reduce_grid(a.n)), dim3(kOptThreads), 0, st, a);
  GCRL_LAUNCHED();
}

__global__ void __launch_bounds__(kOptThreads)
polyak_kernel(float *__restrict__ target, const float *__restrict__ src, int n, float tau, float omt,
              const int *__restrict__ tmap, float *__restrict__ targetT) {
  pdl_wait();
  pdl_launch_dependents();
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const float t = tau * src[e] + omt * target[e];
    target[e] = t;
    if (tmap != nullptr) {
      const int tm = tmap[e];
      if (tm >= 0) targetT[tm] = t;
    }
  }
}

void launch_polyak(float *target, const float *src, int n, float tau, float one_minus_tau,
                   const int *tmap, float *targetT, cudaStream_t st) {
  launch_pdl<PDL_OPTIM>(polyak_kernel, dim3(reduce_grid(n)), dim3(kOptThreads), 0, st, target, src, n, tau, one_minus_tau, tmap, targetT);
  GCRL_LAUNCHED();
}

__global__ void __launch_bounds__(kOptThreads)
sync_transposed_kernel(const float *__restrict__ p, float *__restrict__ pT, const int *__restrict__ tmap, int n) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int tm = tmap[e];
    if (tm >= 0) pT[tm] = p[e];
  }
}

void launch_sync_transposed(const float *p, float *pT, const int *tmap, int n, cudaStream_t st) {
  sync_transposed_kernel<<<reduce_grid(n), kOptThreads, 0, st>>>(p, pT, tmap, n);
  GCRL_LAUNCHED();
}

}  // namespace gcrl
