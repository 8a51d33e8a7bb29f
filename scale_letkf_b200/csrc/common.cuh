// common.cuh -- small device/host helpers shared by every kernel of libletkf_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define LETKF_FULL_MASK 0xffffffffu

namespace letkf {

constexpr int kMaxThreads = 512;   // largest CTA any kernel here launches (das_ns_kernel<13>: 416)
constexpr int kMaxWarps = kMaxThreads / 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(LETKF_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(LETKF_FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(LETKF_FULL_MASK, v, o);
  return v;
}

// Block-wide sum / max; `red` is shared scratch of >= kMaxWarps doubles.  All threads of the
// CTA must call; the result is returned to every thread.  Two barriers.
__device__ __forceinline__ double block_sum(double v, double *red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();   // protect `red` against a previous use
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < nw; ++i) r += red[i];   // same order in every thread: deterministic
  return r;
}
__device__ __forceinline__ double block_max(double v, double *red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = red[0];
  for (int i = 1; i < nw; ++i) r = fmax(r, red[i]);
  return r;
}
__device__ __forceinline__ int block_sum_i(int v, int *red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum_i(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  int r = 0;
  for (int i = 0; i < nw; ++i) r += red[i];
  return r;
}

// Ordered stream compaction step: every thread passes `flag`; returns the exclusive rank of
// this thread among flagged threads and the block total.  `red` >= kMaxWarps ints.
__device__ __forceinline__ int block_rank(bool flag, int *red, int &total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const unsigned m = __ballot_sync(LETKF_FULL_MASK, flag);
  const int r = __popc(m & ((1u << lane) - 1u));
  __syncthreads();
  if (lane == 0) red[w] = __popc(m);
  __syncthreads();
  int base = 0, t = 0;
  for (int i = 0; i < nw; ++i) {
    const int c = red[i];
    if (i < w) base += c;
    t += c;
  }
  total = t;
  return base + r;
}

__host__ __device__ __forceinline__ int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace letkf
