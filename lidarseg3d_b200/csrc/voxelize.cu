// Hard voxelization on the GPU, bit-exact with the reference's sequential numba voxelizer
// (det3d/ops/point_cloud/point_cloud_ops.py:7-55,112-184, called from
//  det3d/datasets/pipelines/segpreprocess.py:148-177), for a batch of frames in one pass.
//
// The sequential rule "voxel id = order of first appearance, keep the first max_points points of a
// voxel in arrival order, stop creating voxels after max_voxels" is reproduced in parallel:
//   1. every point computes floor((p - lo)/vs) with IEEE fp32 sub/div (no reciprocal), packs
//      (frame, z, y, x) into a 64-bit key and inserts it into an open-addressing hash (same-key lanes
//      of a warp are deduplicated with __match_any_sync before touching memory);
//      atomicMin keeps the first point index per voxel, a 5-deep atomicMin chain keeps the
//      max_points smallest point indices per voxel in sorted order;
//   2. an exclusive scan over "is first point of its voxel" flags numbers voxels in first-seen order;
//   3. voxels beyond max_voxels per frame are dropped, rows are filled from the per-voxel index lists.
#include "common.cuh"
#include "scan.cuh"
#include "../../include/ls3d.h"

namespace ls3d {

constexpr int VOX_MAX_FRAMES = 64;
constexpr int VOX_MAX_POINTS = 8;  // points kept per voxel (reference configs use 5)
constexpr long long KEY_EMPTY = -1;

struct VoxParams {
  int n, f, n_frames;
  int off[VOX_MAX_FRAMES + 1];
  float vs[3], lo[3];
  int grid[3];  // x, y, z
  int max_points, max_voxels;
  unsigned long long cap_mask;
};

struct VoxWs {
  long long* keys;   // [cap]
  int* first;        // [cap]
  int* top;          // [cap * max_points]
  int* vox_of_slot;  // [cap]
  int* pslot;        // [n]
  int* rank;         // [n + 1]
  int* block_sums;   // scan scratch
  int* frame_base;   // [n_frames + 1] output row base per frame
  int* frame_rank0;  // [n_frames + 1]
};

__device__ __forceinline__ unsigned long long hash64(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}

__device__ __forceinline__ int frame_of(const VoxParams& p, int i) {
  int f = 0;
  for (int k = 1; k < p.n_frames; ++k)
    if (i >= p.off[k]) f = k;
  return f;
}

__global__ void vox_insert_kernel(const float* __restrict__ pts, VoxParams p, VoxWs w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < p.n;
  long long key = KEY_EMPTY;
  if (in) {
    const float* q = pts + (size_t)i * p.f;
    bool ok = true;
    int c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      // np.floor((p - lo) / vs) in fp32, true division (point_cloud_ops.py:36)
      const float v = floorf(__fdiv_rn(__fsub_rn(q[a], p.lo[a]), p.vs[a]));
      if (!(v >= 0.f) || !(v < (float)p.grid[a])) ok = false;
      c[a] = (int)v;
    }
    if (ok) {
      const int fr = frame_of(p, i);
      key = (((long long)fr * p.grid[2] + c[2]) * p.grid[1] + c[1]) * p.grid[0] + c[0];
    }
  }
  // warp-level dedup: lanes holding the same key elect their lowest lane (= smallest point index)
  const unsigned act = __ballot_sync(0xffffffffu, key != KEY_EMPTY);
  int slot = -1;
  if (key != KEY_EMPTY) {
    const unsigned peers = __match_any_sync(act, key);
    const int leader = __ffs(peers) - 1;
    const int lane = threadIdx.x & 31;
    if (lane == leader) {
      unsigned long long s = hash64((unsigned long long)key) & p.cap_mask;
      while (true) {
        const long long old =
            (long long)atomicCAS((unsigned long long*)&w.keys[s], (unsigned long long)KEY_EMPTY,
                                 (unsigned long long)key);
        if (old == KEY_EMPTY || old == key) break;
        s = (s + 1) & p.cap_mask;
      }
      slot = (int)s;
      atomicMin(&w.first[slot], i);
    }
    slot = __shfl_sync(peers, slot, leader);
    // sorted list of the max_points smallest point indices of the voxel
    int v = i;
    for (int t = 0; t < p.max_points; ++t) {
      const int old = atomicMin(&w.top[(size_t)slot * p.max_points + t], v);
      v = max(old, v);
      if (v == INT_MAX) break;
    }
  }
  if (in) w.pslot[i] = slot;
}

__global__ void vox_frames_kernel(VoxParams p, VoxWs w, int* num_voxels, int* total_out) {
  // rank[] = exclusive scan of first-point flags; rank[n] = total number of voxels before capping
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int base = 0;
    for (int f = 0; f < p.n_frames; ++f) {
      const int r0 = w.rank[p.off[f]];
      const int r1 = w.rank[p.off[f + 1]];
      const int cnt = min(r1 - r0, p.max_voxels);
      w.frame_rank0[f] = r0;
      w.frame_base[f] = base;
      num_voxels[f] = cnt;
      base += cnt;
    }
    w.frame_base[p.n_frames] = base;
    *total_out = base;
  }
}

__global__ void vox_fill_kernel(const float* __restrict__ pts, VoxParams p, VoxWs w, float* voxels,
                                int* coords, int* num_points) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const int slot = w.pslot[i];
  if (slot < 0 || w.first[slot] != i) return;
  const int fr = frame_of(p, i);
  const int local = w.rank[i] - w.frame_rank0[fr];
  if (local >= p.max_voxels) {
    w.vox_of_slot[slot] = -1;
    return;
  }
  const int vid = w.frame_base[fr] + local;
  w.vox_of_slot[slot] = vid;
  long long key = w.keys[slot];
  const int x = (int)(key % p.grid[0]); key /= p.grid[0];
  const int y = (int)(key % p.grid[1]); key /= p.grid[1];
  const int z = (int)(key % p.grid[2]);
  coords[vid * 4 + 0] = fr; coords[vid * 4 + 1] = z; coords[vid * 4 + 2] = y; coords[vid * 4 + 3] = x;
  int cnt = 0;
  float* dst = voxels + (size_t)vid * p.max_points * p.f;
  for (int t = 0; t < p.max_points; ++t) {
    const int src = w.top[(size_t)slot * p.max_points + t];
    if (src != INT_MAX) {
      ++cnt;
      for (int c = 0; c < p.f; ++c) dst[t * p.f + c] = pts[(size_t)src * p.f + c];
    } else {
      for (int c = 0; c < p.f; ++c) dst[t * p.f + c] = 0.f;
    }
  }
  num_points[vid] = cnt;
}

__global__ void vox_point_map_kernel(VoxParams p, VoxWs w, int* point_voxel) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const int slot = w.pslot[i];
  point_voxel[i] = slot < 0 ? -1 : w.vox_of_slot[slot];
}

static __global__ void fill_i32_kernel(int* p, size_t n, int v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = v;
}

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

static unsigned long long hash_capacity(long long n) {
  unsigned long long cap = 1024;
  while (cap < (unsigned long long)(2 * n)) cap <<= 1;
  return cap;
}

static size_t carve(VoxWs* w, uint8_t* base, long long n, int max_points, int n_frames) {
  const unsigned long long cap = hash_capacity(n);
  size_t o = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base ? base + o : nullptr; o += align_up(bytes); return p; };
  w->keys = (long long*)take(cap * 8);
  w->first = (int*)take(cap * 4);
  w->top = (int*)take(cap * 4 * max_points);
  w->vox_of_slot = (int*)take(cap * 4);
  w->pslot = (int*)take((size_t)n * 4);
  w->rank = (int*)take((size_t)(n + 1) * 4);
  w->block_sums = (int*)take(scan_ws_ints(n + 1) * 4);
  w->frame_base = (int*)take((size_t)(n_frames + 1) * 4);
  w->frame_rank0 = (int*)take((size_t)(n_frames + 1) * 4);
  return o;
}

}  // namespace ls3d

extern "C" int ls3d_voxelize_workspace_bytes(int64_t n_points, int32_t max_points, int32_t n_frames,
                                             int64_t* bytes) {
  using namespace ls3d;
  if (!bytes || n_points < 0 || max_points < 1 || max_points > VOX_MAX_POINTS) return LS3D_ERR_ARG;
  VoxWs w;
  *bytes = (int64_t)carve(&w, nullptr, n_points > 0 ? n_points : 1, max_points, n_frames);
  return LS3D_OK;
}

extern "C" int ls3d_voxelize(const float* points, int32_t n_points, int32_t n_feat,
                             const int32_t* frame_off_host, int32_t n_frames, const float* voxel_size,
                             const float* pc_range, int32_t max_points, int32_t max_voxels,
                             void* workspace, int64_t workspace_bytes, float* voxels, int32_t* coords,
                             int32_t* num_points, int32_t* num_voxels, int32_t* total_voxels,
                             int32_t* point_voxel, void* stream) {
  using namespace ls3d;
  if (n_frames < 1 || n_frames > VOX_MAX_FRAMES || max_points < 1 || max_points > VOX_MAX_POINTS ||
      n_feat < 3 || !frame_off_host || !voxel_size || !pc_range || !num_voxels || !total_voxels)
    return LS3D_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_points <= 0) {
    cudaMemsetAsync(num_voxels, 0, sizeof(int) * n_frames, st);
    cudaMemsetAsync(total_voxels, 0, sizeof(int), st);
    return LS3D_OK;
  }
  if (!points || !workspace || !voxels || !coords || !num_points) return LS3D_ERR_ARG;
  VoxParams p;
  p.n = n_points; p.f = n_feat; p.n_frames = n_frames;
  for (int i = 0; i <= n_frames; ++i) p.off[i] = frame_off_host[i];
  if (p.off[0] != 0 || p.off[n_frames] != n_points) return LS3D_ERR_ARG;
  for (int a = 0; a < 3; ++a) {
    p.vs[a] = voxel_size[a];
    p.lo[a] = pc_range[a];
    // grid = round((hi - lo) / vs) in fp32 (point_cloud_ops.py:26-30); rintf = round half to even = np.round
    p.grid[a] = (int)rintf((pc_range[3 + a] - pc_range[a]) / voxel_size[a]);
  }
  p.max_points = max_points; p.max_voxels = max_voxels;
  VoxWs w;
  const size_t need = carve(&w, (uint8_t*)workspace, n_points, max_points, n_frames);
  if ((int64_t)need > workspace_bytes) return LS3D_ERR_ARG;
  const unsigned long long cap = hash_capacity(n_points);
  p.cap_mask = cap - 1;
  cudaMemsetAsync(w.keys, 0xFF, cap * 8, st);
  fill_i32_kernel<<<1024, 256, 0, st>>>(w.first, (size_t)cap, INT_MAX);
  fill_i32_kernel<<<1024, 256, 0, st>>>(w.top, (size_t)cap * max_points, INT_MAX);
  const int T = 256;
  const int G = ls3d_div_up(n_points, T);
  vox_insert_kernel<<<G, T, 0, st>>>(points, p, w);
  // rank[i] = number of first-points with index < i, for i in [0, n]
  {
    const int* pslot = w.pslot;
    const int* first = w.first;
    int* rank = w.rank;
    const int n = n_points;
    auto load = [pslot, first, n] __device__(long long i) -> int {
      if (i >= n) return 0;
      const int s = pslot[i];
      return (s >= 0 && first[s] == (int)i) ? 1 : 0;
    };
    auto store = [rank] __device__(long long i, int v) { rank[i] = v; };
    int e = exclusive_scan(load, store, (long long)n_points + 1, w.block_sums, nullptr, st);
    if (e) return e;
  }
  vox_frames_kernel<<<1, 32, 0, st>>>(p, w, num_voxels, total_voxels);
  vox_fill_kernel<<<G, T, 0, st>>>(points, p, w, voxels, coords, num_points);
  if (point_voxel) vox_point_map_kernel<<<G, T, 0, st>>>(p, w, point_voxel);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}
