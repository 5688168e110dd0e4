// Rulebook ("indice pair") generation for sparse 3-D convolution, output-stationary form.
//
// Replaces spconv 1.x `get_indice_pairs` (external; called through SubMConv3d / SparseConv3d /
// SparseInverseConv3d at reference det3d/models/backbones/scn_unet.py:15-20,39-46,89-160).
// Semantics: kernel offset k = (kz*KY+ky)*KX+kx,
//   out[o] += in[o*s - pad + k] . W[k];  SubM: output sites == input sites (same order);
//   strided: output sites = union of reachable o, numbered by ascending linear index
//   ((b*D+z)*H+y)*W+x;  inverse: the strided conv's pairs with in/out swapped.
//
// Data structure: a dense occupancy bitmap of the level's grid, one uint2 {bits, rank-prefix} per 32
// cells.  Site lookup = one 8-byte read + popc; ascending-linear-index numbering of strided outputs
// falls out of the prefix ranks (no sort, no hash collisions).  A level whose rows are not in linear
// order (level 1: first-seen voxel order) carries a rank -> row permutation.
//
// The tables produced are nbr[k][j] = input row feeding output row j at offset k (or -1): exactly the
// spconv pair list {(k, in, out)} stored by output row, which is what the gather-GEMM consumes.
#include "common.cuh"
#include "scan.cuh"
#include "../../include/ls3d.h"

namespace ls3d {

struct GridDesc {
  const uint2* words;
  const int* perm;  // rank -> row, or nullptr when row == rank
  int B, D, H, W;
};

__device__ __forceinline__ long long cell_of(const GridDesc& g, int b, int z, int y, int x) {
  return (((long long)b * g.D + z) * g.H + y) * g.W + x;
}

// row of the active site at (b,z,y,x) or -1
__device__ __forceinline__ int grid_lookup(const GridDesc& g, int b, int z, int y, int x) {
  if ((unsigned)b >= (unsigned)g.B || (unsigned)z >= (unsigned)g.D || (unsigned)y >= (unsigned)g.H || (unsigned)x >= (unsigned)g.W) return -1;
  const long long c = cell_of(g, b, z, y, x);
  const uint2 w = __ldg(&g.words[c >> 5]);
  const unsigned bit = 1u << (c & 31);
  if (!(w.x & bit)) return -1;
  const int rank = (int)w.y + __popc(w.x & (bit - 1));
  return g.perm ? __ldg(&g.perm[rank]) : rank;
}

__global__ void grid_set_kernel(const int* __restrict__ coords, int m, uint2* words, int B, int D, int H, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int4 c = *reinterpret_cast<const int4*>(coords + (size_t)i * 4);
  // a site outside (B, D, H, W) (example voxelized with another range / shape than the backbone's) is left out of the
  // bitmap - it then takes no part in any rulebook - instead of writing out of bounds
  if ((unsigned)c.x >= (unsigned)B || (unsigned)c.y >= (unsigned)D || (unsigned)c.z >= (unsigned)H || (unsigned)c.w >= (unsigned)W) return;
  const long long cell = (((long long)c.x * D + c.y) * H + c.z) * W + c.w;
  atomicOr(&words[cell >> 5].x, 1u << (cell & 31));
}

__global__ void grid_perm_kernel(const int* __restrict__ coords, int m, const uint2* words, int* perm, int B, int D,
                                 int H, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int4 c = *reinterpret_cast<const int4*>(coords + (size_t)i * 4);
  if ((unsigned)c.x >= (unsigned)B || (unsigned)c.y >= (unsigned)D || (unsigned)c.z >= (unsigned)H || (unsigned)c.w >= (unsigned)W) return;
  const long long cell = (((long long)c.x * D + c.y) * H + c.z) * W + c.w;
  const uint2 w = words[cell >> 5];
  const unsigned bit = 1u << (cell & 31);
  perm[(int)w.y + __popc(w.x & (bit - 1))] = i;
}

struct ConvGeom {
  int k[3], s[3], p[3];
};

// mark every output cell reachable from an active input site
__global__ void grid_mark_strided_kernel(const int* __restrict__ coords, int m, ConvGeom g, uint2* owords, int B, int oD,
                                         int oH, int oW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int4 c = *reinterpret_cast<const int4*>(coords + (size_t)i * 4);
  if ((unsigned)c.x >= (unsigned)B) return;
  for (int kz = 0; kz < g.k[0]; ++kz) {
    const int nz = c.y + g.p[0] - kz;
    if (nz < 0 || nz % g.s[0]) continue;
    const int oz = nz / g.s[0];
    if (oz >= oD) continue;
    for (int ky = 0; ky < g.k[1]; ++ky) {
      const int ny = c.z + g.p[1] - ky;
      if (ny < 0 || ny % g.s[1]) continue;
      const int oy = ny / g.s[1];
      if (oy >= oH) continue;
      for (int kx = 0; kx < g.k[2]; ++kx) {
        const int nx = c.w + g.p[2] - kx;
        if (nx < 0 || nx % g.s[2]) continue;
        const int ox = nx / g.s[2];
        if (ox >= oW) continue;
        const long long cell = (((long long)c.x * oD + oz) * oH + oy) * oW + ox;
        atomicOr(&owords[cell >> 5].x, 1u << (cell & 31));
      }
    }
  }
}

// enumerate active cells in ascending linear order -> coords[rank] = (b,z,y,x)
__global__ void grid_enumerate_kernel(const uint2* __restrict__ words, long long nwords, int D, int H, int W,
                                      int* __restrict__ coords) {
  const long long wi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= nwords) return;
  uint2 w = words[wi];
  unsigned bits = w.x;
  int rank = (int)w.y;
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    long long cell = wi * 32 + b;
    const int x = (int)(cell % W); cell /= W;
    const int y = (int)(cell % H); cell /= H;
    const int z = (int)(cell % D); cell /= D;
    *reinterpret_cast<int4*>(coords + (size_t)rank * 4) = make_int4((int)cell, z, y, x);
    ++rank;
  }
}

// nbr[k][j] = input row at  out_j * s - p + k   (SubM: s = 1, p = K/2)
__global__ void nbr_gather_kernel(GridDesc in, const int* __restrict__ ocoords, int m_out, ConvGeom g,
                                  int* __restrict__ nbr) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int K = g.k[0] * g.k[1] * g.k[2];
  if (t >= (long long)K * m_out) return;
  const int k = (int)(t / m_out);
  const int j = (int)(t % m_out);
  const int kx = k % g.k[2], ky = (k / g.k[2]) % g.k[1], kz = k / (g.k[2] * g.k[1]);
  const int4 c = __ldg(reinterpret_cast<const int4*>(ocoords + (size_t)j * 4));
  nbr[t] = grid_lookup(in, c.x, c.y * g.s[0] - g.p[0] + kz, c.z * g.s[1] - g.p[1] + ky, c.w * g.s[2] - g.p[2] + kx);
}

// inverse conv: nbr_up[k][i] = coarse row o with o*s - p + k == fine_i  (or -1)
__global__ void nbr_scatter_kernel(GridDesc out, const int* __restrict__ icoords, int m_in, ConvGeom g,
                                   int* __restrict__ nbr) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int K = g.k[0] * g.k[1] * g.k[2];
  if (t >= (long long)K * m_in) return;
  const int k = (int)(t / m_in);
  const int i = (int)(t % m_in);
  const int kx = k % g.k[2], ky = (k / g.k[2]) % g.k[1], kz = k / (g.k[2] * g.k[1]);
  const int4 c = __ldg(reinterpret_cast<const int4*>(icoords + (size_t)i * 4));
  const int nz = c.y + g.p[0] - kz, ny = c.z + g.p[1] - ky, nx = c.w + g.p[2] - kx;
  int r = -1;
  if (nz >= 0 && ny >= 0 && nx >= 0 && nz % g.s[0] == 0 && ny % g.s[1] == 0 && nx % g.s[2] == 0)
    r = grid_lookup(out, c.x, nz / g.s[0], ny / g.s[1], nx / g.s[2]);
  nbr[t] = r;
}

static int rank_words(uint2* words, long long nwords, int* block_sums, int* total_out, cudaStream_t st) {
  auto load = [words] __device__(long long i) -> int { return __popc(words[i].x); };
  auto store = [words] __device__(long long i, int v) { words[i].y = (unsigned)v; };
  return exclusive_scan(load, store, nwords, block_sums, total_out, st);
}

static inline long long n_words(long long B, long long D, long long H, long long W) {
  return (B * D * H * W + 31) / 32;
}

}  // namespace ls3d

extern "C" int ls3d_grid_bytes(int32_t B, int32_t D, int32_t H, int32_t W, int64_t* words_bytes,
                               int64_t* scratch_bytes) {
  using namespace ls3d;
  if (!words_bytes || !scratch_bytes || B < 1 || D < 1 || H < 1 || W < 1) return LS3D_ERR_ARG;
  const long long nw = n_words(B, D, H, W);
  *words_bytes = nw * 8;
  *scratch_bytes = (int64_t)scan_ws_ints(nw) * 4;
  return LS3D_OK;
}

extern "C" int ls3d_grid_build(const int32_t* coords, int32_t m, int32_t B, int32_t D, int32_t H, int32_t W,
                               void* words, int32_t* perm, void* scratch, int32_t* total_out, void* stream) {
  using namespace ls3d;
  if (!words || !scratch || (m > 0 && !coords)) return LS3D_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long nw = n_words(B, D, H, W);
  cudaMemsetAsync(words, 0, nw * 8, st);
  if (m > 0) grid_set_kernel<<<ls3d_div_up(m, 256), 256, 0, st>>>(coords, m, (uint2*)words, B, D, H, W);
  int e = rank_words((uint2*)words, nw, (int*)scratch, total_out, st);
  if (e) return e;
  if (perm && m > 0) grid_perm_kernel<<<ls3d_div_up(m, 256), 256, 0, st>>>(coords, m, (const uint2*)words, perm, B, D, H, W);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

static ls3d::ConvGeom make_geom(const int32_t* ksize, const int32_t* stride, const int32_t* pad) {
  ls3d::ConvGeom g;
  for (int a = 0; a < 3; ++a) { g.k[a] = ksize[a]; g.s[a] = stride[a]; g.p[a] = pad[a]; }
  return g;
}

extern "C" int ls3d_grid_build_strided(const int32_t* in_coords, int32_t m_in, int32_t B, const int32_t* ksize,
                                       const int32_t* stride, const int32_t* pad, int32_t oD, int32_t oH,
                                       int32_t oW, void* out_words, void* scratch, int32_t* total_out,
                                       void* stream) {
  using namespace ls3d;
  if (!out_words || !scratch || !total_out || !ksize || !stride || !pad) return LS3D_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long nw = n_words(B, oD, oH, oW);
  cudaMemsetAsync(out_words, 0, nw * 8, st);
  if (m_in > 0)
    grid_mark_strided_kernel<<<ls3d_div_up(m_in, 256), 256, 0, st>>>(in_coords, m_in, make_geom(ksize, stride, pad),
                                                                    (uint2*)out_words, B, oD, oH, oW);
  int e = rank_words((uint2*)out_words, nw, (int*)scratch, total_out, st);
  if (e) return e;
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

extern "C" int ls3d_grid_enumerate(const void* words, int32_t B, int32_t D, int32_t H, int32_t W, int32_t* coords,
                                   void* stream) {
  using namespace ls3d;
  if (!words || !coords) return LS3D_ERR_ARG;
  const long long nw = n_words(B, D, H, W);
  grid_enumerate_kernel<<<ls3d_div_up(nw, 256), 256, 0, (cudaStream_t)stream>>>((const uint2*)words, nw, D, H, W, coords);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

extern "C" int ls3d_rulebook_gather(const void* in_words, const int32_t* in_perm, int32_t B, int32_t D, int32_t H,
                                    int32_t W, const int32_t* out_coords, int32_t m_out, const int32_t* ksize,
                                    const int32_t* stride, const int32_t* pad, int32_t* nbr, void* stream) {
  using namespace ls3d;
  if (m_out <= 0) return LS3D_OK;
  if (!in_words || !out_coords || !nbr) return LS3D_ERR_ARG;
  GridDesc g{(const uint2*)in_words, in_perm, B, D, H, W};
  const ConvGeom cg = make_geom(ksize, stride, pad);
  const long long total = (long long)cg.k[0] * cg.k[1] * cg.k[2] * m_out;
  nbr_gather_kernel<<<ls3d_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(g, out_coords, m_out, cg, nbr);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

extern "C" int ls3d_rulebook_scatter(const void* out_words, int32_t B, int32_t oD, int32_t oH, int32_t oW,
                                     const int32_t* in_coords, int32_t m_in, const int32_t* ksize,
                                     const int32_t* stride, const int32_t* pad, int32_t* nbr, void* stream) {
  using namespace ls3d;
  if (m_in <= 0) return LS3D_OK;
  if (!out_words || !in_coords || !nbr) return LS3D_ERR_ARG;
  GridDesc g{(const uint2*)out_words, nullptr, B, oD, oH, oW};
  const ConvGeom cg = make_geom(ksize, stride, pad);
  const long long total = (long long)cg.k[0] * cg.k[1] * cg.k[2] * m_in;
  nbr_scatter_kernel<<<ls3d_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(g, in_coords, m_in, cg, nbr);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}
