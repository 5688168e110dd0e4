// Exclusive prefix sum over int32 (three-phase: block reduce -> single-block carry scan -> block scan).
// Used for first-seen voxel numbering, bitmap ranks and output-site numbering.
#pragma once
#include "common.cuh"

namespace ls3d {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;                          // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;   // 2048 elements per block

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread (SCAN_THREADS threads); returns exclusive prefix,
// *total = block sum
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
  __shared__ int wsum[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = warp_incl_scan(v, lane);
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = lane < SCAN_THREADS / 32 ? wsum[lane] : 0;
    int si = warp_incl_scan(s, lane);
    if (lane < SCAN_THREADS / 32) wsum[lane] = si - s;
    if (lane == SCAN_THREADS / 32 - 1) *total = si;
  }
  __syncthreads();
  int r = inc - v + wsum[w];
  __syncthreads();
  return r;
}

template <class LoadFn>
__global__ void scan_reduce_kernel(LoadFn load, long long n, int* block_sums) {
  __shared__ int tot;
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) s += load(base + i);
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

// single block: exclusive scan of block_sums[0..nb) in place; writes the grand total to *total_out
static __global__ void scan_carry_kernel(int* block_sums, int nb, int* total_out) {
  __shared__ int tot;
  int carry = 0;
  for (int base = 0; base < nb; base += SCAN_THREADS) {
    const int i = base + threadIdx.x;
    int v = i < nb ? block_sums[i] : 0;
    int ex = block_excl_scan(v, &tot);
    if (i < nb) block_sums[i] = carry + ex;
    carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <class LoadFn, class StoreFn>
__global__ void scan_apply_kernel(LoadFn load, StoreFn store, long long n, const int* block_sums) {
  __shared__ int tot;
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? load(base + i) : 0;
    s += v[i];
  }
  int ex = block_excl_scan(s, &tot) + block_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) store(base + i, ex);
    ex += v[i];
  }
}

static inline size_t scan_ws_ints(long long n) { return (size_t)((n + SCAN_TILE - 1) / SCAN_TILE) + 1; }

// exclusive scan of load(i) for i in [0, n) -> store(i, prefix); *total_out (device) = sum.
template <class LoadFn, class StoreFn>
static inline int exclusive_scan(LoadFn load, StoreFn store, long long n, int* block_sums, int* total_out,
                                 cudaStream_t st) {
  if (n <= 0) {
    if (total_out) cudaMemsetAsync(total_out, 0, sizeof(int), st);
    return 0;
  }
  const int nb = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  scan_reduce_kernel<<<nb, SCAN_THREADS, 0, st>>>(load, n, block_sums);
  scan_carry_kernel<<<1, SCAN_THREADS, 0, st>>>(block_sums, nb, total_out);
  scan_apply_kernel<<<nb, SCAN_THREADS, 0, st>>>(load, store, n, block_sums);
  cudaError_t e = cudaGetLastError();
  return (int)e;
}

}  // namespace ls3d
