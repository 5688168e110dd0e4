// SF-Phase helpers: semantic (class) embedding aggregation and the class-token memory path.
//
//   ls3d_class_embed  : LiDARSemanticFeatureAggregationModule.forward (reference
//                       det3d/models/point_heads/context_module.py:25-53) and
//                       CameraSemanticFeatureAggregationModule.forward
//                       (det3d/models/img_heads/fcn_mseg3d_head.py:24-51): per frame and class,
//                       softmax of the logits over ALL rows (voxels / pixels) of the frame, then
//                       probs[ncls, R] @ feats[R, C].  Deterministic two-level reduction.
//   ls3d_class_tokens : the memory side of SemanticFeatureFusionModule + TransformerDecoderLayer.forward_post
//                       (context_module.py:101-109,211-227,337-339): input projections of the two
//                       embedding sets, then per layer self-attention over the 2*ncls tokens + LayerNorm,
//                       and the k/v projections consumed by the point cross-attention.  The memory never
//                       depends on the points, so all layers run in one launch (one CTA per frame).
#include <cuda_fp16.h>

#include "common.cuh"
#include "../../include/ls3d.h"

namespace ls3d {

constexpr int CE_ROWS = 2048;   // rows per partial block
constexpr int CE_MAXCLS = 32;
constexpr int CE_MAXC = 64;

__device__ __forceinline__ float atomic_max_f(float* addr, float v) {
  // valid for any sign: compare through ordered-int mapping
  int* ai = (int*)addr;
  int old = *ai, assumed;
  do {
    assumed = old;
    if (__int_as_float(assumed) >= v) break;
    old = atomicCAS(ai, assumed, __float_as_int(v));
  } while (assumed != old);
  return __int_as_float(old);
}

static __global__ void fill_f32_kernel(float* p, int n, float v) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = v;
}

__device__ __forceinline__ float ce_ld(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ce_ld(const __half* p) { return __half2float(__ldg(p)); }

// pass 1: per (frame, class) max over rows.  grid = (chunks, frames)
template <typename T>
__global__ void ce_max_kernel(const T* __restrict__ logits, int ld_l, int ncls, const int* __restrict__ seg_off,
                              float* __restrict__ cmax /* [B][ncls] init -inf */) {
  const int b = blockIdx.y;
  const int r0 = seg_off[b], r1 = seg_off[b + 1];
  const int start = r0 + blockIdx.x * CE_ROWS;
  if (start >= r1) return;
  const int end = min(start + CE_ROWS, r1);
  // lane = class (ncls <= 32), warp = row group: a warp reads the ncls contiguous logits of one row per step
  __shared__ float smax8[8][CE_MAXCLS];
  const int c = threadIdx.x & 31, rg = threadIdx.x >> 5;
  float m = -INFINITY;
  if (c < ncls) {
#pragma unroll 4
    for (int r = start + rg; r < end; r += 8) m = fmaxf(m, ce_ld(logits + (size_t)r * ld_l + c));
  }
  smax8[rg][c] = m;
  __syncthreads();
  if (threadIdx.x < ncls) {
    float v = smax8[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 8; ++k) v = fmaxf(v, smax8[k][threadIdx.x]);
    atomic_max_f(&cmax[b * ncls + threadIdx.x], v);
  }
}

// pass 2: partial sums  part[b][chunk][cls][C+1] = sum_r e * [feat | 1]   (probs^T @ [feats | 1] before normalisation)
// Streaming layout: a warp owns rows chunk_start + warp, + 8, ... in groups of CE_GR; lane l owns channels l, l + 32 (, l + 64)
// of [feats | 1] for ALL classes (NCLS4 * 4 x NCH accumulators in registers).  Per group the lanes load the feature rows
// (coalesced, all CE_GR rows in flight), compute the group's e = exp(logit - max) values once into a per-warp shared-memory
// strip and read them back as broadcast float4: per row 2-3 global loads, NCLS4 LDS.128 and ncls * NCH FMAs per lane and no
// block-wide synchronisation inside the row loop.  Cross-warp sum in a fixed order (deterministic).
constexpr int CE_GR = 8;                        // rows per warp step
constexpr int CE_WARPS = 8;
template <typename T, int NCLS4, int NCH>
__global__ void __launch_bounds__(CE_WARPS * 32) ce_partial_kernel(const T* __restrict__ logits, int ld_l, int ncls,
                                                                   const T* __restrict__ feats, int ld_f, int C,
                                                                   const int* __restrict__ seg_off, const float* __restrict__ cmax,
                                                                   float* __restrict__ part, int nchunk) {
  constexpr int NC = NCLS4 * 4;
  const int b = blockIdx.y;
  const int r0 = seg_off[b], r1 = seg_off[b + 1];
  const int start = r0 + blockIdx.x * CE_ROWS;
  float* dst = part + ((size_t)b * nchunk + blockIdx.x) * ncls * (C + 1);
  const int nacc = ncls * (C + 1);
  if (start >= r1) {
    for (int a = threadIdx.x; a < nacc; a += blockDim.x) dst[a] = 0.f;
    return;
  }
  const int end = min(start + CE_ROWS, r1);
  __shared__ __align__(16) float se[CE_WARPS][CE_GR * NC];
  __shared__ float red[CE_WARPS][NCH * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[NC][NCH];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int j = 0; j < NCH; ++j) acc[c][j] = 0.f;
  // this lane's share of a group's e strip: entries lane + 32 i -> (row idx / NC, class idx % NC)
  constexpr int EPL = (CE_GR * NC + 31) / 32;
  float mxv[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    const int c = (lane + 32 * i) % NC;
    mxv[i] = c < ncls ? __ldg(cmax + b * ncls + c) : 0.f;
  }
  float* sew = se[warp];
  for (int g0 = start + warp * CE_GR; g0 < end; g0 += CE_WARPS * CE_GR) {
    float f[CE_GR][NCH];
#pragma unroll
    for (int rr = 0; rr < CE_GR; ++rr) {
      const int r = g0 + rr;
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const int ch = lane + 32 * j;
        f[rr][j] = (r < end) ? (ch < C ? ce_ld(feats + (size_t)r * ld_f + ch) : (ch == C ? 1.f : 0.f)) : 0.f;
      }
    }
    float ev[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const int idx = lane + 32 * i, rr = idx / NC, c = idx % NC, r = g0 + rr;
      ev[i] = (idx < CE_GR * NC && c < ncls && r < end) ? __expf(ce_ld(logits + (size_t)r * ld_l + c) - mxv[i]) : 0.f;
    }
    __syncwarp();                                   // the previous group's strip has been read by every lane
#pragma unroll
    for (int i = 0; i < EPL; ++i)
      if (lane + 32 * i < CE_GR * NC) sew[lane + 32 * i] = ev[i];
    __syncwarp();
#pragma unroll
    for (int rr = 0; rr < CE_GR; ++rr) {
#pragma unroll
      for (int c4 = 0; c4 < NCLS4; ++c4) {
        const float4 e = *reinterpret_cast<const float4*>(sew + rr * NC + c4 * 4);
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          acc[c4 * 4 + 0][j] = fmaf(e.x, f[rr][j], acc[c4 * 4 + 0][j]);
          acc[c4 * 4 + 1][j] = fmaf(e.y, f[rr][j], acc[c4 * 4 + 1][j]);
          acc[c4 * 4 + 2][j] = fmaf(e.z, f[rr][j], acc[c4 * 4 + 2][j]);
          acc[c4 * 4 + 3][j] = fmaf(e.w, f[rr][j], acc[c4 * 4 + 3][j]);
        }
      }
    }
  }
  // cross-warp sum, one class at a time (fixed order: deterministic)
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c >= ncls) break;
#pragma unroll
    for (int j = 0; j < NCH; ++j) red[warp][j * 32 + lane] = acc[c][j];
    __syncthreads();
    if (threadIdx.x <= C) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < CE_WARPS; ++w) v += red[w][threadIdx.x];
      dst[c * (C + 1) + threadIdx.x] = v;
    }
    __syncthreads();
  }
}

// pass 3: emb[b][cls][c] = sum_chunks num / sum_chunks den.  grid = (ncls, frames), thread = channel (C + 1 <= 128):
// a chunk's [C + 1] strip is one coalesced read; sequential over the chunks in a fixed order (deterministic).
__global__ void ce_final_kernel(const float* __restrict__ part, int nchunk, int ncls, int C, float* __restrict__ emb) {
  const int cls = blockIdx.x, b = blockIdx.y, c = threadIdx.x;
  __shared__ float s_den;
  float num = 0.f;
  if (c <= C) {
    const float* src = part + (size_t)b * nchunk * ncls * (C + 1) + (size_t)cls * (C + 1) + c;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int k = 0;
    for (; k + 4 <= nchunk; k += 4) {                  // four independent loads in flight, summed in chunk order
      const float v0 = src[(size_t)(k + 0) * ncls * (C + 1)], v1 = src[(size_t)(k + 1) * ncls * (C + 1)];
      const float v2 = src[(size_t)(k + 2) * ncls * (C + 1)], v3 = src[(size_t)(k + 3) * ncls * (C + 1)];
      a0 += v0; a1 += v1; a2 += v2; a3 += v3;
    }
    for (; k < nchunk; ++k) a0 += src[(size_t)k * ncls * (C + 1)];
    num = (a0 + a1) + (a2 + a3);
  }
  if (c == C) s_den = num;
  __syncthreads();
  if (c < C) emb[((size_t)b * ncls + cls) * C + c] = num / s_den;
}

// ------------------------------------------------------------------------------------------------
// class-token memory path.  Parameters are packed by the host (lidarseg3d_b200/det3d/point_heads.py):
//   head:  P1t[C1][E], b1[E], P2t[C2][E], b2[E]                       (input_proj_embeddings1/2, W^T)
//   layer: Wint[E][3E], bin[3E], Woutt[E][E], bout[E], g1[E], be1[E], Wkt[E][E], bk[E], Wvt[E][E], bv[E]
// Output K, V: [n_layer][B][H][L][dh]  (L = 2*ncls tokens, cam tokens first)
constexpr int CT_MAXL = 48;
constexpr int CT_E = 96;


// L2 prefetch of a parameter block (one 128-byte line per thread and step)
__device__ __forceinline__ void prefetch_l2(const float* p, int nfloats) {
  for (int i = threadIdx.x * 32; i < nfloats; i += blockDim.x * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i));
}

// out[l][e0..e0+3] = bias + sum_c in[l][c] * Wt[c][e0..e0+3]: a thread owns TL tokens x 4 output columns (4 TL accumulators).
// Per 4 input channels it reads 4 weight float4 (global, coalesced over the column groups, next block prefetched) and TL
// activation float4 (shared-memory broadcasts) for 16 TL FMAs.  A column-per-thread mapping spends one broadcast
// LDS.128 per 4 FMAs - on one CTA that is bound by the shared-memory pipe (~4 cycles per warp-wide LDS.128), 12x off the FMA
// rate; with TL = 9 the ratio is 1 : 36.
template <int TL, class StoreFn>
__device__ __forceinline__ void tokens_gemm(const float* __restrict__ Wt, const float* __restrict__ bias, int nin, int nout,
                                             const float* in_s, int ld_in, int L, StoreFn store) {
  const int ncg = nout >> 2;
  const int ntg = (L + TL - 1) / TL;
  for (int t = threadIdx.x; t < ncg * ntg; t += blockDim.x) {
    const int cgp = t % ncg, tg = t / ncg;
    const int e0 = cgp * 4, l0 = tg * TL;
    float4 acc[TL];
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + e0));
#pragma unroll
    for (int i = 0; i < TL; ++i) acc[i] = b4;
    const float* xr[TL];
#pragma unroll
    for (int i = 0; i < TL; ++i) xr[i] = in_s + min(l0 + i, L - 1) * ld_in;
    const float4* wp = reinterpret_cast<const float4*>(Wt + e0);
    float4 wn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wn[i] = __ldg(wp + (size_t)i * ncg);
    for (int c0 = 0; c0 < nin; c0 += 4) {
      float4 w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = wn[i];
      if (c0 + 4 < nin) {
#pragma unroll
        for (int i = 0; i < 4; ++i) wn[i] = __ldg(wp + (size_t)(c0 + 4 + i) * ncg);
      }
#pragma unroll
      for (int i = 0; i < TL; ++i) {
        const float4 x = *reinterpret_cast<const float4*>(xr[i] + c0);
        float4 a = acc[i];
        a.x = fmaf(x.x, w[0].x, a.x); a.y = fmaf(x.x, w[0].y, a.y); a.z = fmaf(x.x, w[0].z, a.z); a.w = fmaf(x.x, w[0].w, a.w);
        a.x = fmaf(x.y, w[1].x, a.x); a.y = fmaf(x.y, w[1].y, a.y); a.z = fmaf(x.y, w[1].z, a.z); a.w = fmaf(x.y, w[1].w, a.w);
        a.x = fmaf(x.z, w[2].x, a.x); a.y = fmaf(x.z, w[2].y, a.y); a.z = fmaf(x.z, w[2].z, a.z); a.w = fmaf(x.z, w[2].w, a.w);
        a.x = fmaf(x.w, w[3].x, a.x); a.y = fmaf(x.w, w[3].y, a.y); a.z = fmaf(x.w, w[3].z, a.z); a.w = fmaf(x.w, w[3].w, a.w);
        acc[i] = a;
      }
    }
#pragma unroll
    for (int i = 0; i < TL; ++i)
      if (l0 + i < L) store(l0 + i, e0, acc[i]);
  }
}

__global__ void __launch_bounds__(512) class_tokens_kernel(const float* __restrict__ emb1, int C1,
                                                           const float* __restrict__ emb2, int C2, int ncls,
                                                           const float* __restrict__ params, int n_layer, int n_head,
                                                           float* __restrict__ Kout, float* __restrict__ Vout,
                                                           float* __restrict__ mem_out) {
  constexpr int E = CT_E;
  const int b = blockIdx.x, B = gridDim.x;
  const int L = 2 * ncls;
  const int dh = E / n_head;
  extern __shared__ float ct_smem[];
  float (*mem)[E] = reinterpret_cast<float (*)[E]>(ct_smem);                       // [L][E]
  // q|k|v rows are 3E + 1 floats apart: the score loop reads k of 32 different tokens in one warp instruction (a 3E stride
  // would put them all in one bank); att starts on a 16-byte boundary again (it is read as float4)
  constexpr int QS = 3 * E + 1;
  const int qkv_floats = (L * QS + 3) & ~3;
  float (*qkv)[QS] = reinterpret_cast<float (*)[QS]>(ct_smem + L * E);             // [L][3E+1]
  float (*att)[E] = reinterpret_cast<float (*)[E]>(ct_smem + L * E + qkv_floats);  // [L][E]
  float* scb = ct_smem + L * E + qkv_floats + L * E;                               // [H][L][L+1]
#define SC(h, i, j) scb[((h) * L + (i)) * (L + 1) + (j)]
  const float* P1t = params;
  const float* b1 = P1t + C1 * E;
  const float* P2t = b1 + E;
  const float* b2 = P2t + C2 * E;
  const float* lp = b2 + E;
  const int layer_sz = E * 3 * E + 3 * E + E * E + E + E + E + E * E + E + E * E + E;
  // every CTA walks the whole parameter block once: ask L2 for all of it up front
  prefetch_l2(params, (C1 + C2 + 2) * E + n_layer * layer_sz);
  // ---- input projections
  // stage the two embedding sets in shared memory (att is free here), then column-owner projections
  float* e1s = &att[0][0];                       // [ncls][C1]
  float* e2s = e1s + ncls * C1;                  // [ncls][C2]   (ncls*(C1+C2) <= L*E)
  for (int o = threadIdx.x; o < ncls * C1; o += blockDim.x) e1s[o] = emb1[(size_t)b * ncls * C1 + o];
  for (int o = threadIdx.x; o < ncls * C2; o += blockDim.x) e2s[o] = emb2[(size_t)b * ncls * C2 + o];
  __syncthreads();
  tokens_gemm<2>(P1t, b1, C1, E, e1s, C1, ncls, [&](int l, int e, float4 v) { *reinterpret_cast<float4*>(&mem[l][e]) = v; });
  tokens_gemm<2>(P2t, b2, C2, E, e2s, C2, ncls, [&](int l, int e, float4 v) { *reinterpret_cast<float4*>(&mem[ncls + l][e]) = v; });
  __syncthreads();
  for (int ly = 0; ly < n_layer; ++ly) {
    const float* Wint = lp + (size_t)ly * layer_sz;
    const float* bin = Wint + E * 3 * E;
    const float* Woutt = bin + 3 * E;
    const float* bout = Woutt + E * E;
    const float* g1 = bout + E;
    const float* be1 = g1 + E;
    const float* Wkt = be1 + E;
    const float* bk = Wkt + E * E;
    const float* Wvt = bk + E;
    const float* bv = Wvt + E * E;
    // in-projection
    tokens_gemm<9>(Wint, bin, E, 3 * E, &mem[0][0], E, L, [&](int l, int e, float4 v) {      // qkv rows are 3E + 1 floats apart
      qkv[l][e] = v.x; qkv[l][e + 1] = v.y; qkv[l][e + 2] = v.z; qkv[l][e + 3] = v.w;
    });
    __syncthreads();
    // scores
    const float scale = rsqrtf((float)dh);
    for (int o = threadIdx.x; o < n_head * L * L; o += blockDim.x) {
      const int h = o / (L * L), i = (o / L) % L, j = o % L;
      float a = 0.f;
      for (int d = 0; d < dh; ++d) a = fmaf(qkv[i][h * dh + d] * scale, qkv[j][E + h * dh + d], a);
      SC(h, i, j) = a;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < n_head * L; o += blockDim.x) {
      const int h = o / L, i = o % L;
      float m = -INFINITY;
      for (int j = 0; j < L; ++j) m = fmaxf(m, SC(h, i, j));
      float s = 0.f;
      for (int j = 0; j < L; ++j) { const float e = __expf(SC(h, i, j) - m); SC(h, i, j) = e; s += e; }
      const float inv = 1.f / s;
      for (int j = 0; j < L; ++j) SC(h, i, j) *= inv;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < L * E; o += blockDim.x) {
      const int i = o / E, e = o % E, h = e / dh;
      float a = 0.f;
      for (int j = 0; j < L; ++j) a = fmaf(SC(h, i, j), qkv[j][2 * E + e], a);
      att[i][e] = a;
    }
    __syncthreads();
    // out-projection + residual (into qkv[:, :E] as scratch)
    tokens_gemm<3>(Woutt, bout, E, E, &att[0][0], E, L, [&](int l, int e, float4 v) {
      qkv[l][e] = mem[l][e] + v.x; qkv[l][e + 1] = mem[l][e + 1] + v.y; qkv[l][e + 2] = mem[l][e + 2] + v.z;
      qkv[l][e + 3] = mem[l][e + 3] + v.w;
    });
    __syncthreads();
    // LayerNorm (norm1), one warp per token
    for (int l = threadIdx.x >> 5; l < L; l += blockDim.x >> 5) {
      const int lane = threadIdx.x & 31;
      float s = 0.f;
      for (int e = lane; e < E; e += 32) s += qkv[l][e];
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s / E;
      float v = 0.f;
      for (int e = lane; e < E; e += 32) { const float d = qkv[l][e] - mean; v = fmaf(d, d, v); }
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const float rstd = rsqrtf(v / E + 1e-5f);
      for (int e = lane; e < E; e += 32) mem[l][e] = (qkv[l][e] - mean) * rstd * g1[e] + be1[e];
    }
    __syncthreads();
    // k / v projections for the point cross-attention of this layer
    {
      const size_t base = ((size_t)ly * B + b) * n_head;
      auto kv_store = [&](float* dst) {
        return [=](int l, int e, float4 v) {                       // [head][token][dh], dh = E / n_head (a multiple of 4)
          const int h = e / dh, d = e - h * dh;
          *reinterpret_cast<float4*>(dst + ((base + h) * L + l) * dh + d) = v;
        };
      };
      tokens_gemm<3>(Wkt, bk, E, E, &mem[0][0], E, L, kv_store(Kout));
      tokens_gemm<3>(Wvt, bv, E, E, &mem[0][0], E, L, kv_store(Vout));
    }
    if (mem_out)
      for (int o = threadIdx.x; o < L * E; o += blockDim.x)
        mem_out[(((size_t)ly * B + b) * L + o / E) * E + o % E] = mem[o / E][o % E];
    __syncthreads();
  }
}

}  // namespace ls3d

extern "C" int ls3d_class_embed_workspace_bytes(int32_t n_frames, int32_t max_rows_per_frame, int32_t ncls, int32_t C,
                                                int64_t* bytes) {
  using namespace ls3d;
  if (!bytes) return LS3D_ERR_ARG;
  const int nchunk = ls3d_div_up(max_rows_per_frame > 0 ? max_rows_per_frame : 1, CE_ROWS);
  *bytes = ((int64_t)n_frames * ncls + (int64_t)n_frames * nchunk * ncls * (C + 1)) * 4;
  return LS3D_OK;
}

extern "C" int ls3d_class_embed(const void* logits, int32_t ld_l, int32_t ncls, const void* feats, int32_t ld_f, int32_t C,
                                int32_t in_fp16, const int32_t* seg_off, int32_t n_frames, int32_t max_rows_per_frame,
                                void* workspace, float* emb, void* stream) {
  using namespace ls3d;
  if (!logits || !feats || !seg_off || !workspace || !emb || ncls > CE_MAXCLS || C > CE_MAXC || ncls < 1)
    return LS3D_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int nchunk = ls3d_div_up(max_rows_per_frame > 0 ? max_rows_per_frame : 1, CE_ROWS);
  float* cmax = (float*)workspace;
  float* part = cmax + (size_t)n_frames * ncls;
  fill_f32_kernel<<<1, 256, 0, st>>>(cmax, n_frames * ncls, -INFINITY);
  dim3 grid(nchunk, n_frames);
  const int ncls4 = (ncls + 3) / 4, nch = (C + 1 + 31) / 32;
#define LS3D_CE_PARTIAL(T, A, B)                                                                                      \
  ce_partial_kernel<T, A, B><<<grid, CE_WARPS * 32, 0, st>>>((const T*)logits, ld_l, ncls, (const T*)feats, ld_f, C, seg_off, \
                                                             cmax, part, nchunk)
#define LS3D_CE_DISPATCH(T)                                                        \
  do {                                                                             \
    if (ncls4 <= 5 && nch <= 2) LS3D_CE_PARTIAL(T, 5, 2);                          \
    else if (ncls4 <= 5) LS3D_CE_PARTIAL(T, 5, 3);                                 \
    else if (nch <= 2) LS3D_CE_PARTIAL(T, 8, 2);                                   \
    else LS3D_CE_PARTIAL(T, 8, 3);                                                 \
  } while (0)
  if (in_fp16) {
    ce_max_kernel<__half><<<grid, 256, 0, st>>>((const __half*)logits, ld_l, ncls, seg_off, cmax);
    LS3D_CE_DISPATCH(__half);
  } else {
    ce_max_kernel<float><<<grid, 256, 0, st>>>((const float*)logits, ld_l, ncls, seg_off, cmax);
    LS3D_CE_DISPATCH(float);
  }
#undef LS3D_CE_DISPATCH
#undef LS3D_CE_PARTIAL
  ce_final_kernel<<<dim3(ncls, n_frames), 128, 0, st>>>(part, nchunk, ncls, C, emb);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

extern "C" int ls3d_class_tokens(const float* emb1, int32_t C1, const float* emb2, int32_t C2, int32_t ncls,
                                 int32_t n_frames, const float* params, int32_t n_layer, int32_t n_head, int32_t d_model,
                                 float* K, float* V, float* mem_out, void* stream) {
  using namespace ls3d;
  if (!emb1 || !emb2 || !params || !K || !V || d_model != CT_E || 2 * ncls > CT_MAXL || n_head < 1 || n_head > 8 ||
      d_model % n_head)
    return LS3D_ERR_ARG;
  const int L = 2 * ncls;
  const size_t smem = ((size_t)L * CT_E * 2 + (((size_t)L * (3 * CT_E + 1) + 3) & ~(size_t)3) + (size_t)n_head * L * (L + 1)) * 4;
  cudaError_t e = cudaFuncSetAttribute(class_tokens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  class_tokens_kernel<<<n_frames, 512, smem, (cudaStream_t)stream>>>(emb1, C1, emb2, C2, ncls, params, n_layer, n_head, K,
                                                                    V, mem_out);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}
