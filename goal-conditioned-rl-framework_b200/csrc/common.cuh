// Shared helpers for the gcrl_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

#include "gcrl_b200.h"

namespace gcrl {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string &m);

#define GCRL_CUDA(expr)                                                                    \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      throw ::gcrl::Error(GCRL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// after every kernel launch: account for it and surface launch-configuration errors
#define GCRL_LAUNCHED()                   \
  do {                                    \
    ::gcrl::count_launch();               \
    GCRL_CUDA(cudaGetLastError());        \
  } while (0)

#define GCRL_REQUIRE(cond, msg)                                          \
  do {                                                                   \
    if (!(cond)) throw ::gcrl::Error(GCRL_ERR_INVALID, std::string(msg)); \
  } while (0)

#define GCRL_API_BEGIN try {
#define GCRL_API_END                                   \
  }                                                    \
  catch (const ::gcrl::Error &e) {                     \
    ::gcrl::set_last_error(e.what());                  \
    return e.code;                                     \
  }                                                    \
  catch (const std::exception &e) {                    \
    ::gcrl::set_last_error(e.what());                  \
    return GCRL_ERR_INVALID;                           \
  }                                                    \
  return GCRL_OK;

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// NVTX range over a host-side phase of the hot path (header-only NVTX3: a no-op unless a profiler is attached).
// The ranges sit on the C-ABI entry points and on the phases inside an update while it is being captured; graph
// replays show up as the enclosing entry-point range.
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
  NvtxRange &operator=(const NvtxRange &) = delete;
};
#define GCRL_NVTX(name) ::gcrl::NvtxRange _gcrl_nvtx_range_(name)

// number of SMs of the current device (148 on B200), cached per process
int sm_count();

// Kernel-launch accounting (gcrl_kernel_launches): every launch site calls count_launch();
// while a stream capture is recording, launches are tallied per graph and credited on replay.
void count_launch(uint64_t n = 1);
uint64_t launch_counter();

template <typename T>
T *dev_alloc(size_t n) {
  T *p = nullptr;
  GCRL_CUDA(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
  return p;
}

// Pinned host staging ring: a slot is reused only after the copy that read it finished.
struct PinnedRing {
  static constexpr int kSlots = 8;
  char *base = nullptr;
  size_t slot_bytes = 0;
  cudaEvent_t ev[kSlots] = {};
  bool used[kSlots] = {};
  int next = 0;
  void init(size_t bytes);
  void destroy();
  // returns a host pointer valid until release(slot, stream) + the stream reaching it
  char *acquire(size_t bytes, int *slot);
  void release(int slot, cudaStream_t st);
};

// ---- programmatic dependent launch -------------------------------------------------------
// Kernels of one update run back to back on one stream (one captured graph).  Launched through launch_pdl, a
// kernel may be scheduled while its predecessor drains; it calls pdl_wait() before touching global memory (the
// predecessor has then completed and flushed) and pdl_launch_dependents() once its own CTAs are running, so the
// launch latency and prologue of kernel N+1 hide under kernel N.  GCRL_NO_PDL=1 turns the attribute off (the two
// device calls are then no-ops).  Measured (profiles/README.md, round 2): worth ~1.5 us per update on the row-slab
// and optimiser kernels; HARMFUL for the weight-gradient launch, whose 576 small CTAs, let in early, all land on
// the 20 SMs the 128-CTA row-slab kernel leaves free (+10 us) -- hence the default class mask.
enum : int { PDL_FUSED = 1, PDL_WGRAD = 2, PDL_OPTIM = 4, PDL_OTHER = 8, PDL_P2P = 16 };   // GCRL_PDL_MASK selects kernel classes
bool pdl_enabled(int cls = PDL_OTHER);
template <int CLS = PDL_OTHER, typename... KArgs, typename... Args>
void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled(CLS) ? 1 : 0;
  GCRL_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}

// ---- device helpers -------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Exact unsigned division by a runtime constant (Granlund-Montgomery, 32-bit n < 2^31).
struct FastDiv {
  uint32_t d, mul, shr;
  __host__ __device__ FastDiv() : d(1), mul(0), shr(0) {}
  __host__ __device__ explicit FastDiv(uint32_t div) : d(div) {
    if (div == 1) { mul = 0; shr = 0; return; }
    uint32_t l = 0;
    while ((1u << l) < div) ++l;              // ceil(log2(d))
    shr = l - 1;
    mul = (uint32_t)(((1ull << 32) * ((1ull << l) - div)) / div + 1);
  }
  __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
    if (d == 1) return n;
    uint32_t t = __umulhi(n, mul);
    return (t + ((n - t) >> 1)) >> shr;
#else
    return n / d;
#endif
  }
};

__device__ __forceinline__ float4 ldg_stream4(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream4(float4 *p, const float4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

}  // namespace gcrl
