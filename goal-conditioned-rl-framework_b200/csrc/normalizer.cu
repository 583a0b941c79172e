// Running mean/std normaliser on the device (sm_100a).
//
// Replaces RunningNormalizer (reference src/utils.py:68-117): float64 running
// (mean, var, count) merged with the batch moments by Chan's parallel-variance formula,
// normalize = clip((x - mean) / (sqrt(var) + 1e-8), +-clip).
//
// Kernels
//   norm_update_seq_kernel : batches of <= 8192 rows (the reference's regime: 2-4 rows per env
//                         per step, src/env.py:165-172).  One thread per column walks the rows in
//                         order, accumulating in the INPUT dtype with unfused IEEE operations --
//                         exactly what numpy's axis-0 np.mean / np.var do for dim >= 2 -- then
//                         applies _update_from_moments in float64 without FMA contraction, so the
//                         running statistics are bit-identical to the reference's.
//   norm_partial_kernel : one warp per column, lanes stride over the rows of the CTA's row
//                         slab with a per-lane Welford accumulator, lanes merged with
//                         shuffles (Chan merge of (n, mean, M2) triples) -> one partial per
//                         (CTA, column).  Bound: HBM read of 4|8 * n * dim bytes.
//   norm_merge_kernel   : fixed-order merge of the CTA partials -> batch (mean, var, n),
//                         then the reference's _update_from_moments (src/utils.py:82-94).
//   norm_apply_kernel   : element-wise normalise + clip, float64 math, f64 or f32 output.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace gcrl {

struct Moments {
  double n, mean, m2;
};

__device__ __forceinline__ Moments chan_merge(const Moments &a, const Moments &b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  Moments r;
  r.n = a.n + b.n;
  const double delta = b.mean - a.mean;
  r.mean = a.mean + delta * (b.n / r.n);
  r.m2 = a.m2 + b.m2 + delta * delta * (a.n * b.n / r.n);
  return r;
}

__device__ __forceinline__ double shfl_xor_d(double v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}

constexpr int kNormWarps = 8;
constexpr int kNormThreads = kNormWarps * 32;

// Batch moments of one row block, HBM-bound: the block's elements are read ONCE, in flat (row-major) order, by
// T = dim * floor(256 / dim) threads -- a multiple of the row width, so thread t sees column t % dim only and
// consecutive threads read consecutive addresses.  Every thread keeps float64 sums of (x - p) and (x - p)^2 about
// a per-column pivot p (the block's first row: shifted sums do not cancel), four independent loads in flight;
// the threads of a column are then merged in thread order with Chan's formula (fixed order: deterministic).
template <typename T>
__global__ void __launch_bounds__(kNormThreads)
norm_partial_kernel(const T *__restrict__ x, int64_t n, int dim, int64_t rows_per_block,
                    Moments *__restrict__ partials /* [gridDim.x][dim] */) {
  __shared__ Moments sm[kNormThreads];
  const int tid = threadIdx.x;
  const int per_col = kNormThreads / dim;            // threads per column
  const int T_ = per_col * dim;                      // active threads
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(n, r0 + rows_per_block);
  Moments acc{0.0, 0.0, 0.0};
  if (tid < T_ && r0 < r1) {
    const int c = tid % dim;
    const double p = double(x[r0 * dim + c]);
    const T *base = x + r0 * dim;
    const int64_t cnt = (r1 - r0) * dim;
    double s1 = 0.0, s2 = 0.0, k = 0.0;
    int64_t e = tid;
    for (; e + 3 * int64_t(T_) < cnt; e += 4 * int64_t(T_)) {
      const double v0 = double(base[e]) - p, v1 = double(base[e + T_]) - p, v2 = double(base[e + 2 * int64_t(T_)]) - p,
                   v3 = double(base[e + 3 * int64_t(T_)]) - p;
      s1 += (v0 + v1) + (v2 + v3);
      s2 += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
      k += 4.0;
    }
    for (; e < cnt; e += T_) {
      const double v = double(base[e]) - p;
      s1 += v;
      s2 += v * v;
      k += 1.0;
    }
    if (k > 0.0) {
      acc.n = k;
      acc.mean = p + s1 / k;
      acc.m2 = fmax(s2 - s1 * s1 / k, 0.0);
    }
  }
  sm[tid] = acc;
  __syncthreads();
  if (tid < dim) {                                   // column tid: its threads tid, tid + dim, ... in order
    Moments m = sm[tid];
    for (int j = 1; j < per_col; ++j) m = chan_merge(m, sm[tid + j * dim]);
    partials[size_t(blockIdx.x) * dim + tid] = m;
  }
}

struct NormState {
  double *mean, *var, *count;  // device, [dim], [dim], [1]
};

// unfused IEEE arithmetic in the input dtype (numpy never contracts a*b+c)
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// RunningNormalizer._update_from_moments (src/utils.py:82-94), float64, operation for operation:
//   total = count + n;  delta = bmean - mean;  mean += delta * n / total
//   m2 = var * count + m_b + delta^2 * count * n / total;  var = m2 / total
// m_b = batch_var * n is formed by the CALLER: numpy evaluates it in the batch dtype (a float32
// batch variance times a Python int stays float32, src/utils.py:88).
__device__ __forceinline__ void update_from_moments(NormState st, int c, double count, double bmean,
                                                    double m_b, double n) {
  const double tot = __dadd_rn(count, n);
  const double mean = st.mean[c];
  const double delta = __dsub_rn(bmean, mean);
  const double new_mean = __dadd_rn(mean, __ddiv_rn(__dmul_rn(delta, n), tot));
  const double m2 = __dadd_rn(__dadd_rn(__dmul_rn(st.var[c], count), m_b),
                              __ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(delta, delta), count), n), tot));
  st.mean[c] = new_mean;
  st.var[c] = __ddiv_rn(m2, tot);
  if (c == 0) *st.count = tot;
}

constexpr int64_t kSeqMaxRows = 8192;

template <typename T>
__global__ void __launch_bounds__(128)
norm_update_seq_kernel(const T *__restrict__ x, int n, int dim, NormState st) {
  const int c = threadIdx.x;  // single CTA, one thread per column (coalesced across the row)
  const double count = *st.count;
  __syncthreads();            // every column reads the old count before thread 0 replaces it
  if (c >= dim) return;
  T s = x[c];
  for (int r = 1; r < n; ++r) s = add_rn(s, x[size_t(r) * dim + c]);
  const T mean = div_rn(s, T(n));
  T d = sub_rn(x[c], mean);
  T v = mul_rn(d, d);
  for (int r = 1; r < n; ++r) {
    d = sub_rn(x[size_t(r) * dim + c], mean);
    v = add_rn(v, mul_rn(d, d));
  }
  const T var = div_rn(v, T(n));
  update_from_moments(st, c, count, double(mean), double(mul_rn(var, T(n))), double(n));
}

// All row-block partials of column c folded by ONE WARP: lane l takes blocks l, l + 32, ... (their loads are
// independent and issued together -- a single thread walking 592 partials paid one memory round trip each,
// 200 us), then a shuffle tree with the lower lane first.  The tree's shape depends on nblocks only, so every
// rank of a data-parallel run folds identically.
__device__ __forceinline__ Moments merge_partials_warp(const Moments *__restrict__ partials, int nblocks, int dim, int c,
                                                       int lane) {
  Moments acc{0.0, 0.0, 0.0};
  for (int i0 = lane; i0 < nblocks; i0 += 128) {
    Moments v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + 32 * u;
      v[u] = i < nblocks ? partials[size_t(i) * dim + c] : Moments{0.0, 0.0, 0.0};
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc = chan_merge(acc, v[u]);
  }
#pragma unroll
  for (int m = 1; m <= 16; m <<= 1) {
    Moments o{shfl_xor_d(acc.n, m), shfl_xor_d(acc.mean, m), shfl_xor_d(acc.m2, m)};
    acc = (lane & m) ? chan_merge(o, acc) : chan_merge(acc, o);
  }
  return acc;
}

constexpr int kMergeThreads = 1024;

__global__ void __launch_bounds__(kMergeThreads)
norm_merge_kernel(const Moments *__restrict__ partials, int nblocks, int dim, NormState st) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double count = *st.count;
  __syncthreads();            // every column reads the old count before column 0 replaces it
  for (int c = warp; c < dim; c += kMergeThreads / 32) {
    const Moments b = merge_partials_warp(partials, nblocks, dim, c, lane);
    if (lane == 0) update_from_moments(st, c, count, b.mean, b.m2, b.n);
  }
}

// batch moments only (no state update): what a rank contributes to a data-parallel update
__global__ void __launch_bounds__(kMergeThreads)
norm_collect_kernel(const Moments *__restrict__ partials, int nblocks, int dim, Moments *__restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < dim; c += kMergeThreads / 32) {
    const Moments b = merge_partials_warp(partials, nblocks, dim, c, lane);
    if (lane == 0) out[c] = b;
  }
}

// normalize(x) = clip((x - mean) / (sqrt(var) + 1e-8), +-clip) in float64, operation for operation as NumPy does it
// (src/utils.py:96-98).  The per-column mean and denominator sit in shared memory; four independent loads per
// thread and iteration keep enough bytes in flight for HBM (one load per thread ran at 23 % of it).
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
norm_apply_kernel(const TI *__restrict__ x, int64_t n, int dim, NormState st, double clip,
                  TO *__restrict__ out, int64_t out_stride, int64_t out_col0, FastDiv ddiv) {
  __shared__ double s_mean[128], s_den[128];
  if (int(threadIdx.x) < dim) {
    s_mean[threadIdx.x] = st.mean[threadIdx.x];
    s_den[threadIdx.x] = sqrt(st.var[threadIdx.x]) + 1e-8;
  }
  __syncthreads();
  const int64_t total = n * dim;
  const bool small = total < (int64_t(1) << 31);
  const int64_t step = int64_t(gridDim.x) * blockDim.x;
  auto emit = [&](int64_t e, TI xv) {
    const int64_t r = small ? int64_t(ddiv.div(uint32_t(e))) : e / dim;
    const int c = int(e - r * dim);
    double z = (double(xv) - s_mean[c]) / s_den[c];
    z = fmin(fmax(z, -clip), clip);
    out[r * out_stride + out_col0 + c] = TO(z);
  };
  int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; e + 3 * step < total; e += 4 * step) {
    const TI v0 = x[e], v1 = x[e + step], v2 = x[e + 2 * step], v3 = x[e + 3 * step];
    emit(e, v0); emit(e + step, v1); emit(e + 2 * step, v2); emit(e + 3 * step, v3);
  }
  for (; e < total; e += step) emit(e, x[e]);
}

}  // namespace gcrl

using namespace gcrl;

struct gcrl_norm {
  int device = 0, dim = 0;
  double clip = 5.0;
  NormState st{};
  double *d_state = nullptr;  // mean[dim] | var[dim] | count[2]
  Moments *d_partials = nullptr;
  int max_blocks = 0;
  void *d_x = nullptr;
  size_t x_cap = 0;
  double *d_out = nullptr;
  size_t out_cap = 0;
  PinnedRing stage;
};

static void norm_update_device(gcrl_norm *h, const void *x_dev, int64_t n, int is_f64,
                               cudaStream_t st) {
  if (n <= 0) return;
  GCRL_REQUIRE(h->dim <= 128, "normaliser dim > 128 not supported");
  if (n <= kSeqMaxRows && h->dim >= 2) {   // reference-scale batch: bit-exact numpy order
    if (is_f64)
      norm_update_seq_kernel<double><<<1, 128, 0, st>>>(static_cast<const double *>(x_dev), int(n), h->dim, h->st);
    else
      norm_update_seq_kernel<float><<<1, 128, 0, st>>>(static_cast<const float *>(x_dev), int(n), h->dim, h->st);
    GCRL_LAUNCHED();
    return;
  }
  const int64_t rows_per_block = std::max<int64_t>(256, (n + h->max_blocks - 1) / h->max_blocks);
  const int nblocks = int((n + rows_per_block - 1) / rows_per_block);
  if (is_f64)
    norm_partial_kernel<double><<<nblocks, kNormWarps * 32, 0, st>>>(
        static_cast<const double *>(x_dev), n, h->dim, rows_per_block, h->d_partials);
  else
    norm_partial_kernel<float><<<nblocks, kNormWarps * 32, 0, st>>>(
        static_cast<const float *>(x_dev), n, h->dim, rows_per_block, h->d_partials);
  GCRL_LAUNCHED();
  norm_merge_kernel<<<1, kMergeThreads, 0, st>>>(h->d_partials, nblocks, h->dim, h->st);
  GCRL_LAUNCHED();
}

static const void *norm_stage_in(gcrl_norm *h, const void *x_host, int64_t n, int is_f64,
                                 cudaStream_t st) {
  const size_t bytes = size_t(n) * h->dim * (is_f64 ? 8 : 4);
  if (bytes > h->x_cap) {
    GCRL_CUDA(cudaStreamSynchronize(st));
    if (h->d_x) GCRL_CUDA(cudaFree(h->d_x));
    h->x_cap = std::max<size_t>(bytes * 2, 4096);
    h->d_x = dev_alloc<char>(h->x_cap);
  }
  int slot;
  char *p = h->stage.acquire(bytes, &slot);
  std::memcpy(p, x_host, bytes);
  GCRL_CUDA(cudaMemcpyAsync(h->d_x, p, bytes, cudaMemcpyHostToDevice, st));
  h->stage.release(slot, st);
  return h->d_x;
}

template <typename TO>
static void norm_apply_launch(gcrl_norm *h, const void *x_dev, int64_t n, int is_f64, TO *out,
                              int64_t stride, int64_t col0, cudaStream_t st) {
  if (n <= 0) return;
  const int64_t total = n * h->dim;
  const int blocks = int(std::min<int64_t>((total + 255) / 256, int64_t(sm_count()) * 8));
  const FastDiv dd{uint32_t(h->dim)};
  if (is_f64)
    norm_apply_kernel<double, TO><<<blocks, 256, 0, st>>>(static_cast<const double *>(x_dev), n,
                                                         h->dim, h->st, h->clip, out, stride, col0, dd);
  else
    norm_apply_kernel<float, TO><<<blocks, 256, 0, st>>>(static_cast<const float *>(x_dev), n,
                                                        h->dim, h->st, h->clip, out, stride, col0, dd);
  GCRL_LAUNCHED();
}

extern "C" {

int gcrl_norm_create(gcrl_norm **out, int device, int dim, double clip_range, double eps_count) {
  GCRL_API_BEGIN
  GCRL_REQUIRE(out != nullptr && dim >= 1 && dim <= 128, "need 1 <= dim <= 128");
  GCRL_CUDA(cudaSetDevice(device));
  auto *h = new gcrl_norm();
  try {
    h->device = device;
    h->dim = dim;
    h->clip = clip_range;
    h->d_state = dev_alloc<double>(size_t(2 * dim + 2));
    h->st.mean = h->d_state;
    h->st.var = h->d_state + dim;
    h->st.count = h->d_state + 2 * dim;
    h->max_blocks = sm_count() * 4;
    h->d_partials = dev_alloc<Moments>(size_t(h->max_blocks) * dim);
    h->stage.init(size_t(1) << 16);
  } catch (...) {
    delete h;
    throw;
  }
  *out = h;
  std::vector<double> init(size_t(2 * dim + 2), 0.0);
  for (int i = 0; i < dim; ++i) init[dim + i] = 1.0;  // var = 1, mean = 0 (src/utils.py:70-71)
  init[2 * dim] = init[2 * dim + 1] = eps_count;      // count = eps (src/utils.py:72)
  GCRL_CUDA(cudaMemcpy(h->d_state, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
  GCRL_API_END
}

int gcrl_norm_destroy(gcrl_norm *h) {
  GCRL_API_BEGIN
  if (h == nullptr) return GCRL_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  cudaFree(h->d_state); cudaFree(h->d_partials);
  if (h->d_x) cudaFree(h->d_x);
  if (h->d_out) cudaFree(h->d_out);
  h->stage.destroy();
  delete h;
  GCRL_API_END
}

int gcrl_norm_update(gcrl_norm *h, const void *x_host, int64_t n, int is_f64, void *stream) {
  GCRL_API_BEGIN
  GCRL_NVTX("gcrl_norm_update");
  GCRL_REQUIRE(h != nullptr && (x_host != nullptr || n == 0) && n >= 0, "bad arguments");
  GCRL_CUDA(cudaSetDevice(h->device));
  if (n == 0) return GCRL_OK;
  cudaStream_t st = as_stream(stream);
  const void *xd = norm_stage_in(h, x_host, n, is_f64, st);
  norm_update_device(h, xd, n, is_f64, st);
  GCRL_API_END
}

int gcrl_norm_batch_moments(gcrl_norm *h, const void *x_host, int64_t n, int is_f64, double *moments_host,
                            void *stream) {
  GCRL_API_BEGIN
  GCRL_REQUIRE(h != nullptr && moments_host != nullptr && (x_host != nullptr || n == 0) && n >= 0, "bad arguments");
  GCRL_REQUIRE(h->dim <= 128, "normaliser dim > 128 not supported");
  GCRL_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    std::memset(moments_host, 0, size_t(h->dim) * sizeof(Moments));
    return GCRL_OK;
  }
  const void *xd = norm_stage_in(h, x_host, n, is_f64, st);
  const int64_t rows_per_block = std::max<int64_t>(256, (n + h->max_blocks - 2) / (h->max_blocks - 1));
  const int nblocks = int((n + rows_per_block - 1) / rows_per_block);
  if (is_f64)
    norm_partial_kernel<double><<<nblocks, kNormWarps * 32, 0, st>>>(static_cast<const double *>(xd), n, h->dim,
                                                                     rows_per_block, h->d_partials);
  else
    norm_partial_kernel<float><<<nblocks, kNormWarps * 32, 0, st>>>(static_cast<const float *>(xd), n, h->dim,
                                                                    rows_per_block, h->d_partials);
  GCRL_LAUNCHED();
  Moments *out = h->d_partials + size_t(h->max_blocks - 1) * h->dim;      // last slab of the scratch
  norm_collect_kernel<<<1, kMergeThreads, 0, st>>>(h->d_partials, nblocks, h->dim, out);
  GCRL_LAUNCHED();
  GCRL_CUDA(cudaMemcpyAsync(moments_host, out, size_t(h->dim) * sizeof(Moments), cudaMemcpyDeviceToHost, st));
  GCRL_CUDA(cudaStreamSynchronize(st));
  GCRL_API_END
}

int gcrl_norm_update_moments(gcrl_norm *h, const double *moments_host, int parts, void *stream) {
  GCRL_API_BEGIN
  GCRL_REQUIRE(h != nullptr && moments_host != nullptr && parts >= 1, "bad arguments");
  GCRL_REQUIRE(parts <= h->max_blocks, "too many parts");
  GCRL_REQUIRE(h->dim <= 128, "normaliser dim > 128 not supported");
  GCRL_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = as_stream(stream);
  const size_t bytes = size_t(parts) * h->dim * sizeof(Moments);
  int slot;
  char *p = h->stage.acquire(bytes, &slot);
  std::memcpy(p, moments_host, bytes);
  GCRL_CUDA(cudaMemcpyAsync(h->d_partials, p, bytes, cudaMemcpyHostToDevice, st));
  h->stage.release(slot, st);
  norm_merge_kernel<<<1, kMergeThreads, 0, st>>>(h->d_partials, parts, h->dim, h->st);
  GCRL_LAUNCHED();
  GCRL_API_END
}

int gcrl_norm_update_dev(gcrl_norm *h, const void *x_dev, int64_t n, int is_f64, void *stream) {
  GCRL_API_BEGIN
  GCRL_NVTX("gcrl_norm_update_dev");
  GCRL_REQUIRE(h != nullptr && (x_dev != nullptr || n == 0) && n >= 0, "bad arguments");
  GCRL_CUDA(cudaSetDevice(h->device));
  norm_update_device(h, x_dev, n, is_f64, as_stream(stream));
  GCRL_API_END
}

int gcrl_norm_apply(gcrl_norm *h, const void *x_host, int64_t n, int is_f64, double *out_host,
                    void *stream) {
  GCRL_API_BEGIN
  GCRL_NVTX("gcrl_norm_apply");
  GCRL_REQUIRE(h != nullptr && n >= 0 && (n == 0 || (x_host && out_host)), "bad arguments");
  GCRL_CUDA(cudaSetDevice(h->device));
  if (n == 0) return GCRL_OK;
  cudaStream_t st = as_stream(stream);
  const void *xd = norm_stage_in(h, x_host, n, is_f64, st);
  const size_t cnt = size_t(n) * h->dim;
  if (cnt > h->out_cap) {
    GCRL_CUDA(cudaStreamSynchronize(st));
    if (h->d_out) GCRL_CUDA(cudaFree(h->d_out));
    h->out_cap = cnt * 2;
    h->d_out = dev_alloc<double>(h->out_cap);
  }
  norm_apply_launch<double>(h, xd, n, is_f64, h->d_out, h->dim, 0, st);
  GCRL_CUDA(cudaMemcpyAsync(out_host, h->d_out, cnt * 8, cudaMemcpyDeviceToHost, st));
  GCRL_CUDA(cudaStreamSynchronize(st));
  GCRL_API_END
}

int gcrl_norm_apply_dev_f32(gcrl_norm *h, const void *x_dev, int64_t n, int is_f64,
                            float *out_dev, int64_t out_stride, int64_t out_col0, void *stream) {
  GCRL_API_BEGIN
  GCRL_NVTX("gcrl_norm_apply_dev_f32");
  GCRL_REQUIRE(h != nullptr && n >= 0 && (n == 0 || (x_dev && out_dev)), "bad arguments");
  GCRL_REQUIRE(out_stride >= h->dim + out_col0 && out_col0 >= 0, "bad output stride / column");
  GCRL_CUDA(cudaSetDevice(h->device));
  norm_apply_launch<float>(h, x_dev, n, is_f64, out_dev, out_stride, out_col0, as_stream(stream));
  GCRL_API_END
}

int gcrl_norm_get_state(gcrl_norm *h, double *mean, double *var, double *count,
                        double *clip_range, void *stream) {
  GCRL_API_BEGIN
  GCRL_REQUIRE(h != nullptr, "handle is NULL");
  GCRL_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = as_stream(stream);
  std::vector<double> host(size_t(2 * h->dim + 2));
  GCRL_CUDA(cudaMemcpyAsync(host.data(), h->d_state, host.size() * 8, cudaMemcpyDeviceToHost, st));
  GCRL_CUDA(cudaStreamSynchronize(st));
  if (mean) std::memcpy(mean, host.data(), size_t(h->dim) * 8);
  if (var) std::memcpy(var, host.data() + h->dim, size_t(h->dim) * 8);
  if (count) *count = host[size_t(2 * h->dim)];
  if (clip_range) *clip_range = h->clip;
  GCRL_API_END
}

int gcrl_norm_set_state(gcrl_norm *h, const double *mean, const double *var, double count,
                        double clip_range, void *stream) {
  GCRL_API_BEGIN
  GCRL_REQUIRE(h != nullptr && mean && var, "bad arguments");
  GCRL_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = as_stream(stream);
  std::vector<double> host(size_t(2 * h->dim + 2));
  std::memcpy(host.data(), mean, size_t(h->dim) * 8);
  std::memcpy(host.data() + h->dim, var, size_t(h->dim) * 8);
  host[size_t(2 * h->dim)] = host[size_t(2 * h->dim + 1)] = count;
  h->clip = clip_range;
  GCRL_CUDA(cudaMemcpyAsync(h->d_state, host.data(), host.size() * 8, cudaMemcpyHostToDevice, st));
  GCRL_CUDA(cudaStreamSynchronize(st));
  GCRL_API_END
}

}  // extern "C"
