// Devoxelization: exact 3-nearest voxel centres per raw point + inverse-distance interpolation.
//
// Replaces three_nn_kernel_fast / three_interpolate_kernel_fast
// (reference det3d/ops/pointnet2_batch/src/interpolate_gpu.cu:16-59,84-104) as driven by
// three_interpolate_wrap (det3d/models/point_heads/point_utils.py:8-52).
//
// The reference scans all M voxel centres per point (O(N*M)).  Voxel centres live on the level-1
// lattice, so the same answer is found by scanning the occupancy bitmap in growing boxes around the
// point's own cell until the third-best distance is provably smaller than anything outside the box;
// points that do not converge (far outside the range / isolated) go to an exact brute-force pass.
// Semantics kept bit-exact with the restated kernel: d = (dx*dx + dy*dy) + dz*dz in fp32 without
// FMA contraction, candidates ordered by (d, voxel row) which equals the sequential strict-'<' scan;
// centre = (idx + 0.5) * voxel_size + range_min with separate fp32 mul / add
// (det3d/core/utils/common_utils.py:74-90).
#include "common.cuh"
#include "../../include/ls3d.h"

namespace ls3d {

struct NNParams {
  const uint2* words;
  const int* perm;        // rank -> voxel row (level-1 rows are in first-seen order)
  int B, D, H, W;         // level-1 grid (z, y, x extents of the bitmap)
  float vs[3], lo[3];     // x, y, z
  int n;
  int n_frames;
  const int* point_off;   // [n_frames + 1] device
  const int* voxel_off;   // [n_frames + 1] device
};

struct Top3 {
  float d[3];
  int i[3];
  __device__ __forceinline__ void init() {
    d[0] = d[1] = d[2] = INFINITY;
    i[0] = i[1] = i[2] = 0;
  }
  // insert keeping (d, index) lexicographic order == sequential strict '<' cascade
  __device__ __forceinline__ void push(float dd, int idx) {
    if (dd < d[0] || (dd == d[0] && idx < i[0] && d[0] != INFINITY)) {
      d[2] = d[1]; i[2] = i[1]; d[1] = d[0]; i[1] = i[0]; d[0] = dd; i[0] = idx;
    } else if (dd < d[1] || (dd == d[1] && idx < i[1] && d[1] != INFINITY)) {
      d[2] = d[1]; i[2] = i[1]; d[1] = dd; i[1] = idx;
    } else if (dd < d[2] || (dd == d[2] && idx < i[2] && d[2] != INFINITY)) {
      d[2] = dd; i[2] = idx;
    }
  }
};

__device__ __forceinline__ float centre(int idx, float vs, float lo) {
  return __fadd_rn(__fmul_rn((float)idx + 0.5f, vs), lo);
}

__device__ __forceinline__ float dist2(float ux, float uy, float uz, float x, float y, float z) {
  const float dx = __fsub_rn(ux, x), dy = __fsub_rn(uy, y), dz = __fsub_rn(uz, z);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ int frame_of_point(const int* off, int nf, int i) {
  int f = 0;
  for (int k = 1; k < nf; ++k)
    if (i >= __ldg(off + k)) f = k;
  return f;
}

// Box of "radius" r around cell c: the per-axis half-widths are scaled so that the box is roughly a metric cube (a z cell is
// twice as tall as it is wide in the shipped configs, so it needs half as many z layers for the same lower bound):
// ra[a] = ceil(r * min(vs) / vs[a]).  Any box shape is exact - termination only uses the true distance to the first cells
// outside the scanned box.
struct Box {
  int ra[3];
};
__device__ __forceinline__ Box make_box(const NNParams& p, int r) {
  const float vmin = fminf(p.vs[0], fminf(p.vs[1], p.vs[2]));
  Box b;
#pragma unroll
  for (int a = 0; a < 3; ++a) b.ra[a] = max(1, (int)ceilf((float)r * vmin / p.vs[a] - 1e-4f));
  return b;
}
// lower bound of the distance from u to any centre outside the box (INFINITY: the box covers the whole frame grid)
__device__ __forceinline__ float box_bound(const NNParams& p, const int* c, const float* u, const Box& bx) {
  const int G[3] = {p.W, p.H, p.D};
  float bound = INFINITY;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int lo_i = c[a] - bx.ra[a] - 1, hi_i = c[a] + bx.ra[a] + 1;
    if (lo_i >= 0) bound = fminf(bound, u[a] - centre(lo_i, p.vs[a], p.lo[a]));
    if (hi_i <= G[a] - 1) bound = fminf(bound, centre(hi_i, p.vs[a], p.lo[a]) - u[a]);
  }
  return bound;
}
__device__ __forceinline__ bool box_done(float bound, float d3) {
  return bound == INFINITY || (bound > 0.f && d3 < bound * bound * 0.9999f);
}
__device__ __forceinline__ void point_cell(const NNParams& p, const float* u, int* c) {
  const int G[3] = {p.W, p.H, p.D};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float t = floorf((u[a] - p.lo[a]) / p.vs[a]);
    t = fminf(fmaxf(t, 0.f), (float)(G[a] - 1));
    if (!(t == t)) t = 0.f;
    c[a] = (int)t;
  }
}
// one bitmap row (fixed z, y; x in [x0, x1]) of frame f into `best`; cells with |x - cx| <= skip_x are skipped when the row
// itself lies inside the previously scanned box (inner_row)
__device__ __forceinline__ void scan_row(const NNParams& p, int f, int z, int y, int x0, int x1, bool inner_row, int cx,
                                         int skip_x, float ux, float uy, float uz, Top3& best) {
  const float cz = centre(z, p.vs[2], p.lo[2]);
  const float cy = centre(y, p.vs[1], p.lo[1]);
  const long long base = (((long long)f * p.D + z) * p.H + y) * p.W;
  const long long w0 = (base + x0) >> 5, w1 = (base + x1) >> 5;
  for (long long wi = w0; wi <= w1; ++wi) {
    const uint2 w = __ldg(&p.words[wi]);
    unsigned bits = w.x;
    if (!bits) continue;
    const long long cell0 = wi << 5;
    const long long lo_c = base + x0, hi_c = base + x1;
    if (cell0 < lo_c) bits &= ~0u << (int)(lo_c - cell0);
    if (cell0 + 31 > hi_c) bits &= ~0u >> (int)(cell0 + 31 - hi_c);
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      const int x = (int)(cell0 + b - base);
      if (inner_row && abs(x - cx) <= skip_x) continue;  // visited with the previous box
      const float dd = dist2(ux, uy, uz, centre(x, p.vs[0], p.lo[0]), cy, cz);
      if (dd <= best.d[2]) {
        const int rank = (int)w.y + __popc(w.x & ((1u << b) - 1u));
        best.push(dd, __ldg(&p.perm[rank]));
      }
    }
  }
}

// pass 1: one thread per point, boxes of radius 2 and 4 (most points of a LiDAR sweep resolve here)
__global__ void three_nn_grid_kernel(const float* __restrict__ pts, int ld_p, NNParams p, float* __restrict__ dist2_out,
                                     int* __restrict__ idx_out, int* __restrict__ todo, int* __restrict__ todo_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const int f = frame_of_point(p.point_off, p.n_frames, i);
  const float ux = pts[(size_t)i * ld_p + 1], uy = pts[(size_t)i * ld_p + 2], uz = pts[(size_t)i * ld_p + 3];
  const float u[3] = {ux, uy, uz};
  int c[3];
  point_cell(p, u, c);
  Top3 best;
  best.init();
  bool done = false;
  Box prev;
  prev.ra[0] = prev.ra[1] = prev.ra[2] = -1;
  for (int r = 2; r <= 4 && !done; r <<= 1) {
    const Box bx = make_box(p, r);
    const int z0 = max(c[2] - bx.ra[2], 0), z1 = min(c[2] + bx.ra[2], p.D - 1);
    const int y0 = max(c[1] - bx.ra[1], 0), y1 = min(c[1] + bx.ra[1], p.H - 1);
    const int x0 = max(c[0] - bx.ra[0], 0), x1 = min(c[0] + bx.ra[0], p.W - 1);
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) {
        const bool inner_row = (prev.ra[0] >= 0) && (abs(z - c[2]) <= prev.ra[2]) && (abs(y - c[1]) <= prev.ra[1]);
        scan_row(p, f, z, y, x0, x1, inner_row, c[0], prev.ra[0], ux, uy, uz, best);
      }
    prev = bx;
    done = box_done(box_bound(p, c, u, bx), best.d[2]);
  }
  if (!done) {
    todo[atomicAdd(todo_count, 1)] = i;
    return;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    dist2_out[(size_t)i * 3 + k] = best.d[k];
    idx_out[(size_t)i * 3 + k] = best.i[k];
  }
}

// pass 2: one WARP per unresolved point, boxes of radius 8, 16, 32 rescanned from scratch with the rows dealt to the lanes;
// the lanes' candidate lists are merged with (d, row)-ordered pushes, which makes the result independent of the order
// (== the sequential strict-'<' scan).  Points that still do not converge are compacted into todo2 for the brute-force pass.
__global__ void three_nn_grid_warp_kernel(const float* __restrict__ pts, int ld_p, NNParams p, const int* __restrict__ todo,
                                          const int* __restrict__ todo_count, float* __restrict__ dist2_out,
                                          int* __restrict__ idx_out, int* __restrict__ todo2, int* __restrict__ todo2_count) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int cnt = *todo_count;
  for (int t = wid; t < cnt; t += nw) {
    const int i = todo[t];
    const int f = frame_of_point(p.point_off, p.n_frames, i);
    const float ux = pts[(size_t)i * ld_p + 1], uy = pts[(size_t)i * ld_p + 2], uz = pts[(size_t)i * ld_p + 3];
    const float u[3] = {ux, uy, uz};
    int c[3];
    point_cell(p, u, c);
    bool done = false;
    Top3 best;
    for (int r = 8; r <= 32 && !done; r <<= 1) {
      const Box bx = make_box(p, r);
      const int z0 = max(c[2] - bx.ra[2], 0), z1 = min(c[2] + bx.ra[2], p.D - 1);
      const int y0 = max(c[1] - bx.ra[1], 0), y1 = min(c[1] + bx.ra[1], p.H - 1);
      const int x0 = max(c[0] - bx.ra[0], 0), x1 = min(c[0] + bx.ra[0], p.W - 1);
      const int ny = y1 - y0 + 1, nrows = (z1 - z0 + 1) * ny;
      best.init();
      for (int q = lane; q < nrows; q += 32) scan_row(p, f, z0 + q / ny, y0 + q % ny, x0, x1, false, 0, 0, ux, uy, uz, best);
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        float od[3];
        int oi[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          od[k] = __shfl_xor_sync(0xffffffffu, best.d[k], o);
          oi[k] = __shfl_xor_sync(0xffffffffu, best.i[k], o);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (od[k] != INFINITY && od[k] <= best.d[2]) best.push(od[k], oi[k]);
      }
      done = box_done(box_bound(p, c, u, bx), best.d[2]);
    }
    if (lane == 0) {
      if (done) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          dist2_out[(size_t)i * 3 + k] = best.d[k];
          idx_out[(size_t)i * 3 + k] = best.i[k];
        }
      } else {
        todo2[atomicAdd(todo2_count, 1)] = i;
      }
    }
  }
}

// exact fallback: one block per unresolved point, all voxels of its frame
__global__ void three_nn_brute_kernel(const float* __restrict__ pts, int ld_p, NNParams p,
                                      const int* __restrict__ vcoords, const int* __restrict__ todo,
                                      const int* __restrict__ todo_count, float* __restrict__ dist2_out,
                                      int* __restrict__ idx_out) {
  __shared__ float sd[256 * 3];
  __shared__ int si[256 * 3];
  const int cnt = *todo_count;
  for (int t = blockIdx.x; t < cnt; t += gridDim.x) {
    const int i = todo[t];
    const int f = frame_of_point(p.point_off, p.n_frames, i);
    const float ux = pts[(size_t)i * ld_p + 1], uy = pts[(size_t)i * ld_p + 2], uz = pts[(size_t)i * ld_p + 3];
    const int v0 = p.voxel_off[f], v1 = p.voxel_off[f + 1];
    Top3 best;
    best.init();
    for (int v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
      const int4 c = __ldg(reinterpret_cast<const int4*>(vcoords + (size_t)v * 4));
      const float dd = dist2(ux, uy, uz, centre(c.w, p.vs[0], p.lo[0]), centre(c.z, p.vs[1], p.lo[1]),
                             centre(c.y, p.vs[2], p.lo[2]));
      if (dd <= best.d[2]) best.push(dd, v);
    }
    for (int k = 0; k < 3; ++k) {
      sd[threadIdx.x * 3 + k] = best.d[k];
      si[threadIdx.x * 3 + k] = best.i[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      Top3 m;
      m.init();
      for (int q = 0; q < blockDim.x * 3; ++q)
        if (sd[q] != INFINITY && sd[q] <= m.d[2]) m.push(sd[q], si[q]);
      for (int k = 0; k < 3; ++k) {
        dist2_out[(size_t)i * 3 + k] = m.d[k];
        idx_out[(size_t)i * 3 + k] = m.i[k];
      }
    }
    __syncthreads();
  }
}

// out[n, :C] = sum_k w_k * feat[idx[n,k], :C],  w_k = (1/(sqrt(d2_k)+1e-8)) / sum_k(...)
// (point_utils.py:29-32 + interpolate_gpu.cu:84-104), features row-major [M, ld_f]
__global__ void three_interpolate_kernel(const float* __restrict__ feat, int ld_f, int C, const float* __restrict__ d2,
                                         const int* __restrict__ idx, int n, float* __restrict__ out, int ld_out, int rnd) {
  const int c4 = C / 4;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * c4) return;
  const int i = (int)(t / c4);
  const int c = (int)(t % c4) * 4;
  float w[3];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    w[k] = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(d2 + (size_t)i * 3 + k)), 1e-8f));
  }
  s = __fadd_rn(__fadd_rn(w[0], w[1]), w[2]);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float wk = __fdiv_rn(w[k], s);
    const float4 v = ldg_f4(feat + (size_t)__ldg(idx + (size_t)i * 3 + k) * ld_f + c);
    if (k == 0) {
      acc.x = __fmul_rn(wk, v.x); acc.y = __fmul_rn(wk, v.y); acc.z = __fmul_rn(wk, v.z); acc.w = __fmul_rn(wk, v.w);
    } else {
      acc.x = __fadd_rn(acc.x, __fmul_rn(wk, v.x)); acc.y = __fadd_rn(acc.y, __fmul_rn(wk, v.y));
      acc.z = __fadd_rn(acc.z, __fmul_rn(wk, v.z)); acc.w = __fadd_rn(acc.w, __fmul_rn(wk, v.w));
    }
  }
  if (rnd) { acc.x = to_tf32(acc.x); acc.y = to_tf32(acc.y); acc.z = to_tf32(acc.z); acc.w = to_tf32(acc.w); }
  *reinterpret_cast<float4*>(out + (size_t)i * ld_out + c) = acc;
}

// Reference-signature 3-NN (any point sets, no lattice): the sequential strict-'<' scan of three_nn_kernel_fast with the
// known points staged through shared memory in 256-point tiles (every thread of the block walks the same tile).
__global__ void three_nn_tiled_kernel(int n, int m, const float* __restrict__ unknown, const float* __restrict__ known,
                                      float* __restrict__ dist2_out, int* __restrict__ idx_out) {
  __shared__ float sk[256 * 3];
  const int b = blockIdx.y;
  unknown += (size_t)b * n * 3;
  known += (size_t)b * m * 3;
  dist2_out += (size_t)b * n * 3;
  idx_out += (size_t)b * n * 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  const float ux = live ? unknown[(size_t)i * 3] : 0.f, uy = live ? unknown[(size_t)i * 3 + 1] : 0.f,
              uz = live ? unknown[(size_t)i * 3 + 2] : 0.f;
  // the reference initialises best = 1e40 (inf as float) and idx = 0 and inserts on strict '<' only
  float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int k0 = 0; k0 < m; k0 += 256) {
    const int cnt = min(256, m - k0);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * 3; t += blockDim.x) sk[t] = known[(size_t)k0 * 3 + t];
    __syncthreads();
    for (int k = 0; k < cnt; ++k) {
      const float d = dist2(ux, uy, uz, sk[k * 3], sk[k * 3 + 1], sk[k * 3 + 2]);
      if (d < b1) {
        b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k0 + k;
      } else if (d < b2) {
        b3 = b2; i3 = i2; b2 = d; i2 = k0 + k;
      } else if (d < b3) {
        b3 = d; i3 = k0 + k;
      }
    }
  }
  if (live) {
    dist2_out[(size_t)i * 3] = b1; dist2_out[(size_t)i * 3 + 1] = b2; dist2_out[(size_t)i * 3 + 2] = b3;
    idx_out[(size_t)i * 3] = i1; idx_out[(size_t)i * 3 + 1] = i2; idx_out[(size_t)i * 3 + 2] = i3;
  }
}

// Row offsets of the frames of a batch-sorted tensor: off[b] = first row whose batch column is >= b (off[B] = n).  The rows of
// `points` / `voxel_coords` are concatenated frame by frame by the reference's collate (collate.py:141-150), so this is the
// segment table every per-frame op of the heads needs (the reference masks `batch_idx == i` per frame: point_utils.py:19-21).
template <typename T>
__global__ void frame_offsets_kernel(const T* __restrict__ col, long long stride, int n, int B, int* __restrict__ off) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > B) return;
  int lo = 0, hi = n;                     // lower bound of b in the non-decreasing batch column
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((float)col[(size_t)mid * stride] < (float)b) lo = mid + 1;
    else hi = mid;
  }
  off[b] = lo;
}

}  // namespace ls3d

extern "C" int ls3d_frame_offsets(const void* batch_col, int32_t is_float, int64_t stride, int32_t n, int32_t n_frames,
                                  int32_t* off, void* stream) {
  using namespace ls3d;
  if (!off || n < 0 || n_frames < 1 || (n > 0 && !batch_col) || stride < 1) return LS3D_ERR_ARG;
  const int threads = 32, blocks = (n_frames + 1 + threads - 1) / threads;
  if (is_float)
    frame_offsets_kernel<float><<<blocks, threads, 0, (cudaStream_t)stream>>>((const float*)batch_col, stride, n, n_frames, off);
  else
    frame_offsets_kernel<int><<<blocks, threads, 0, (cudaStream_t)stream>>>((const int*)batch_col, stride, n, n_frames, off);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

namespace ls3d {
}  // namespace ls3d

// drop-in for three_nn_wrapper_fast(b, n, m, unknown, known, dist2, idx) (pointnet2_api.cpp:10-24, interpolate.cpp:17-30)
extern "C" int ls3d_three_nn(int32_t b, int32_t n, int32_t m, const float* unknown, const float* known, float* dist2,
                             int32_t* idx, void* stream) {
  using namespace ls3d;
  if (b <= 0 || n <= 0) return LS3D_OK;
  if (!unknown || !known || !dist2 || !idx || m < 0) return LS3D_ERR_ARG;
  dim3 grid(ls3d_div_up(n, 256), b);
  three_nn_tiled_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, unknown, known, dist2, idx);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

extern "C" int ls3d_three_nn_grid(const float* points, int32_t ld_p, int32_t n, const void* words, const int32_t* perm,
                                  int32_t B, int32_t D, int32_t H, int32_t W, const float* voxel_size_xyz,
                                  const float* range_min_xyz, const int32_t* point_off, const int32_t* voxel_off,
                                  const int32_t* voxel_coords, int32_t* todo, int32_t* todo_count, float* dist2,
                                  int32_t* idx, void* stream) {
  using namespace ls3d;
  if (n <= 0) return LS3D_OK;
  if (!points || !words || !perm || !point_off || !voxel_off || !voxel_coords || !todo || !todo_count || !dist2 || !idx)
    return LS3D_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  NNParams p;
  p.words = (const uint2*)words; p.perm = perm; p.B = B; p.D = D; p.H = H; p.W = W;
  for (int a = 0; a < 3; ++a) { p.vs[a] = voxel_size_xyz[a]; p.lo[a] = range_min_xyz[a]; }
  p.n = n; p.n_frames = B; p.point_off = point_off; p.voxel_off = voxel_off;
  // todo_count[0]: points left by pass 1 (list todo[0, n)), todo_count[1]: points left by pass 2 (list todo[n, 2n))
  cudaMemsetAsync(todo_count, 0, 2 * sizeof(int), st);
  three_nn_grid_kernel<<<ls3d_div_up(n, 128), 128, 0, st>>>(points, ld_p, p, dist2, idx, todo, todo_count);
  three_nn_grid_warp_kernel<<<4 * ls3d_num_sms(), 256, 0, st>>>(points, ld_p, p, todo, todo_count, dist2, idx, todo + n,
                                                                todo_count + 1);
  three_nn_brute_kernel<<<592, 256, 0, st>>>(points, ld_p, p, voxel_coords, todo + n, todo_count + 1, dist2, idx);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

extern "C" int ls3d_three_interpolate(const float* feat, int32_t ld_f, int32_t C, const float* dist2, const int32_t* idx,
                                      int32_t n, float* out, int32_t ld_out, int32_t round_out, void* stream) {
  using namespace ls3d;
  if (n <= 0) return LS3D_OK;
  if (!feat || !dist2 || !idx || !out || (C & 3) || (ld_f & 3) || (ld_out & 3)) return LS3D_ERR_ARG;
  three_interpolate_kernel<<<ls3d_div_up((long long)n * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(
      feat, ld_f, C, dist2, idx, n, out, ld_out, round_out);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}
