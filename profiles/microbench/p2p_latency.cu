// Peer-memory micro-benchmark (2 GPUs, one process): how long does a one-shot read of a gradient-sized buffer
// (0.55 MB) from the peer take, and what is the flag ping-pong round trip?  Build: nvcc -arch=sm_100a -O3.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void read_peer(const float4 *peer, const float4 *local, float4 *out, int n4) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += gridDim.x * blockDim.x) {
    const float4 a = __ldcv(peer + q), b = __ldcv(local + q);
    out[q] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}
// ping-pong: GPU a writes k to peer flag, waits for own flag == k
__global__ void pingpong(volatile unsigned *mine, volatile unsigned *theirs, int iters, int first, long long *cycles) {
  const long long t0 = clock64();
  for (int k = 1; k <= iters; ++k) {
    if (first) { *theirs = k; __threadfence_system(); while (*mine < (unsigned)k) {} }
    else { while (*mine < (unsigned)k) {} *theirs = k; __threadfence_system(); }
  }
  *cycles = clock64() - t0;
}
int main() {
  int n; CK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("need 2 GPUs\n"); return 0; }
  int can01, can10; CK(cudaDeviceCanAccessPeer(&can01, 0, 1)); CK(cudaDeviceCanAccessPeer(&can10, 1, 0));
  printf("peer access 0->1 %d, 1->0 %d\n", can01, can10);
  const int n4 = 138496 / 4 * 1;   // one network's flat gradient in 16-byte groups (~0.55 MB)
  float4 *b0, *b1, *o0; unsigned *f0, *f1; long long *c0, *c1;
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaMalloc(&b0, n4 * 16)); CK(cudaMalloc(&o0, n4 * 16)); CK(cudaMalloc(&f0, 4)); CK(cudaMalloc(&c0, 8));
  CK(cudaMemset(b0, 0, n4 * 16)); CK(cudaMemset(f0, 0, 4));
  CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0)); CK(cudaMalloc(&b1, n4 * 16)); CK(cudaMalloc(&f1, 4)); CK(cudaMalloc(&c1, 8));
  CK(cudaMemset(b1, 0, n4 * 16)); CK(cudaMemset(f1, 0, 4));
  CK(cudaDeviceSynchronize());
  CK(cudaSetDevice(0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int grid : {34, 68, 136, 272}) {
    for (int i = 0; i < 3; ++i) read_peer<<<grid, 256>>>(b1, b0, o0, n4);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 50; ++i) read_peer<<<grid, 256>>>(b1, b0, o0, n4);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("one-shot peer read of %.2f MB, grid %3d: %.2f us per launch (incl. ~2 us launch)\n", n4 * 16 / 1e6, grid, ms * 1000 / 50);
  }
  for (int i = 0; i < 3; ++i) read_peer<<<136, 256>>>(b0, b0, o0, n4);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 50; ++i) read_peer<<<136, 256>>>(b0, b0, o0, n4);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("same kernel on local memory: %.2f us per launch\n", ms * 1000 / 50);
  const int iters = 2000;
  CK(cudaSetDevice(1)); pingpong<<<1, 1>>>(f1, f0, iters, 0, c1);
  CK(cudaSetDevice(0)); pingpong<<<1, 1>>>(f0, f1, iters, 1, c0);
  CK(cudaDeviceSynchronize()); CK(cudaSetDevice(1)); CK(cudaDeviceSynchronize());
  long long cyc; CK(cudaMemcpy(&cyc, c1, 8, cudaMemcpyDeviceToHost));
  int khz; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 1));
  printf("flag ping-pong: %.2f us per round trip (%lld cycles / %d iters at %d kHz nominal)\n", cyc / double(iters) / (khz * 1e-3), cyc, iters, khz);
  return 0;
}
