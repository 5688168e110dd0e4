// Fused epilogue of the gather-GEMM kernels (shared by gather_gemm.cu and gather_gemm_bf16x3.cu).
#pragma once
#include "common.cuh"
#include "../../include/ls3d.h"

namespace ls3d {

constexpr int TILE_M = 128;
constexpr int KCH = 32;                 // channels per K chunk
constexpr int MAX_KOFF = 27;
constexpr int MAX_TOK = 48;
constexpr int DHEAD = 24;
constexpr int MASK_RING = 32;           // > total ring depth: producers are never more than that many tiles ahead of the MMA warp

__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t chunk) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4);
}

// ------------------------------------------------------------------------------------------------------------------
// Epilogue of one 128-row tile (4 warps, thread `et` owns tile row `et`; its accumulator row starts at TMEM address trow).
// Global traffic is coalesced: residual rows come in and results go out through a [128 x 16]-column shared-memory panel
// (a warp request covers whole 64-byte row segments); per-column vectors (folded BN scale/shift, LayerNorm gamma/beta)
// live in shared memory.  Finished pre-LayerNorm values are parked in TMEM over the accumulator columns.
// ------------------------------------------------------------------------------------------------------------------
constexpr int PANEL = 16;
constexpr int STG_LD = PANEL + 4;      // floats; 80-byte rows: 16-byte aligned, conflict-free for row-per-thread float4
constexpr int COLV = 256;              // stride of the per-column vectors in smem: scale, shift, g0, b0, g1, b1

__device__ __forceinline__ void bar_sync_epilogue() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

// Each epilogue warp stages only its own 32 rows, so a __syncwarp() is all the synchronisation a panel needs.
// warp-cooperative [<=32 rows x w] panel copy global -> stgw (zero outside); rows0 = first global row of the warp
__device__ __forceinline__ void panel_load(float* stgw, const float* src, int ld, int rows0, int rows_valid, int c0, int w,
                                            int lane) {
  if (((ld | c0) & 3) == 0 && (w & 3) == 0) {
#pragma unroll
    for (int i = 0; i < PANEL / 4; ++i) {
      const int idx = i * 32 + lane, rr = idx / (PANEL / 4), cn = (idx % (PANEL / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rr < rows_valid && cn < w) v = ldg_f4(src + (size_t)(rows0 + rr) * ld + c0 + cn);
      *reinterpret_cast<float4*>(stgw + rr * STG_LD + cn) = v;
    }
  } else {
    for (int idx = lane; idx < 32 * PANEL; idx += 32) {
      const int rr = idx / PANEL, cn = idx % PANEL;
      stgw[rr * STG_LD + cn] = (rr < rows_valid && cn < w) ? __ldg(src + (size_t)(rows0 + rr) * ld + c0 + cn) : 0.f;
    }
  }
}
// warp-cooperative [<=32 rows x w] panel copy stgw -> global: 8 rows x 64 contiguous bytes per warp request
__device__ __forceinline__ void panel_store(const float* stgw, float* dst, int ld, int rows0, int rows_valid, int c0, int w,
                                             int lane) {
  if (((ld | c0) & 3) == 0 && (w & 3) == 0) {
#pragma unroll
    for (int i = 0; i < PANEL / 4; ++i) {
      const int idx = i * 32 + lane, rr = idx / (PANEL / 4), cn = (idx % (PANEL / 4)) * 4;
      if (rr < rows_valid && cn < w)
        *reinterpret_cast<float4*>(dst + (size_t)(rows0 + rr) * ld + c0 + cn) = *reinterpret_cast<const float4*>(stgw + rr * STG_LD + cn);
    }
  } else {
    for (int idx = lane; idx < 32 * PANEL; idx += 32) {
      const int rr = idx / PANEL, cn = idx % PANEL;
      if (rr < rows_valid && cn < w) dst[(size_t)(rows0 + rr) * ld + c0 + cn] = stgw[rr * STG_LD + cn];
    }
  }
}

// dual > 0: the accumulator is the sum of two partial tiles, columns [0, n) and [dual, dual + n) (bf16x3 stacked-W mode).
__device__ __forceinline__ void epilogue_tile(const ls3d_gemm_args& p, const uint32_t trow, const int tile_row0, const int et,
                                              const float* colv, float* stg_all, const uint32_t dual = 0) {
  const int lane = et & 31;
  const int row0 = tile_row0 + (et & ~31);             // first global row of this warp's 32-row slice
  const int r = tile_row0 + et;
  const bool live = r < p.m_out;
  const int rows_valid = max(0, min(32, p.m_out - row0));
  auto rnd = [&](float x) -> float { return p.round_out ? to_tf32(x) : x; };
  float* stg = stg_all + (et & ~31) * STG_LD;          // this warp's staging rows
  float* my = stg + lane * STG_LD;

  if (p.epi == LS3D_EPI_ATTN) {
    // q = acc + bias ; per head softmax(q.K^T * scale) V over the frame's class tokens
    int f = 0;
    for (int i = 1; i < p.n_frames; ++i)
      if (r >= p.frame_off[i]) f = i;
    const int L = p.n_tok;
    for (int h = 0; h < p.n_head; ++h) {
      uint32_t raw[24];
      tmem_ld8(trow + h * DHEAD, raw);
      tmem_ld8(trow + h * DHEAD + 8, raw + 8);
      tmem_ld8(trow + h * DHEAD + 16, raw + 16);
      tmem_ld_wait();
      float q[DHEAD];
#pragma unroll
      for (int d = 0; d < DHEAD; ++d) q[d] = __uint_as_float(raw[d]) + colv[COLV + h * DHEAD + d];
      if (dual) {
        tmem_ld8(trow + dual + h * DHEAD, raw);
        tmem_ld8(trow + dual + h * DHEAD + 8, raw + 8);
        tmem_ld8(trow + dual + h * DHEAD + 16, raw + 16);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < DHEAD; ++d) q[d] += __uint_as_float(raw[d]);
      }
      const float* kh = p.attn_k + ((size_t)(f * p.n_head + h) * L) * DHEAD;
      const float* vh = p.attn_v + ((size_t)(f * p.n_head + h) * L) * DHEAD;
      float sc[MAX_TOK];
      float mx = -INFINITY;
#pragma unroll
      for (int l = 0; l < MAX_TOK; ++l) {
        if (l < L) {
          float a = 0.f;
#pragma unroll
          for (int d4 = 0; d4 < DHEAD / 4; ++d4) {
            float4 kv = ldg_f4(kh + l * DHEAD + d4 * 4);
            a = fmaf(q[d4 * 4 + 0], kv.x, a);
            a = fmaf(q[d4 * 4 + 1], kv.y, a);
            a = fmaf(q[d4 * 4 + 2], kv.z, a);
            a = fmaf(q[d4 * 4 + 3], kv.w, a);
          }
          a *= p.attn_scale;
          sc[l] = a;
          mx = fmaxf(mx, a);
        }
      }
      float den = 0.f;
      float o[DHEAD];
#pragma unroll
      for (int d = 0; d < DHEAD; ++d) o[d] = 0.f;
#pragma unroll
      for (int l = 0; l < MAX_TOK; ++l) {
        if (l < L) {
          float e = __expf(sc[l] - mx);
          den += e;
#pragma unroll
          for (int d4 = 0; d4 < DHEAD / 4; ++d4) {
            float4 vv = ldg_f4(vh + l * DHEAD + d4 * 4);
            o[d4 * 4 + 0] = fmaf(e, vv.x, o[d4 * 4 + 0]);
            o[d4 * 4 + 1] = fmaf(e, vv.y, o[d4 * 4 + 1]);
            o[d4 * 4 + 2] = fmaf(e, vv.z, o[d4 * 4 + 2]);
            o[d4 * 4 + 3] = fmaf(e, vv.w, o[d4 * 4 + 3]);
          }
        }
      }
      const float inv = 1.f / den;
      // the head's 24 outputs leave through two panels (16 + 8 columns)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int w = half ? DHEAD - PANEL : PANEL;
#pragma unroll
        for (int d4 = 0; d4 < PANEL / 4; ++d4) {
          const int d = half * PANEL + d4 * 4;
          if (d < DHEAD)
            *reinterpret_cast<float4*>(my + d4 * 4) =
                make_float4(rnd(o[d] * inv), rnd(o[d + 1] * inv), rnd(o[d + 2] * inv), rnd(o[d + 3] * inv));
        }
        __syncwarp();
        panel_store(stg, p.out, p.ld_out, row0, rows_valid, h * DHEAD + half * PANEL, w, lane);
        __syncwarp();
      }
    }
    return;
  }

  const bool masked = p.row_mask && live && (__ldg(p.row_mask + (size_t)r * p.ld_mask) != 1.0f);
  const bool has_affine = p.scale || p.shift;
  // one 16-column panel of finished values into v[]: acc * scale + shift, residual, ReLU, channel-reduction add.
  // Every option is a warp-uniform branch around a short unrolled pass over the register panel.
  auto finished_panel = [&](int c0, float* v) {
    const int w = min(PANEL, p.cout - c0);
    uint32_t raw[PANEL];
    tmem_ld16(trow + c0, raw);
    float rv[PANEL];
    if (p.res_mode) {
      panel_load(stg, p.res, p.ld_res, row0, rows_valid, c0, w, lane);
      __syncwarp();
#pragma unroll
      for (int j4 = 0; j4 < PANEL / 4; ++j4) {
        const float4 t = *reinterpret_cast<const float4*>(my + j4 * 4);
        rv[j4 * 4] = t.x; rv[j4 * 4 + 1] = t.y; rv[j4 * 4 + 2] = t.z; rv[j4 * 4 + 3] = t.w;
      }
      __syncwarp();
    }
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < PANEL; ++j) v[j] = __uint_as_float(raw[j]);
    if (dual) {
      tmem_ld16(trow + dual + c0, raw);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < PANEL; ++j) v[j] += __uint_as_float(raw[j]);
    }
    if (has_affine) {
#pragma unroll
      for (int j4 = 0; j4 < PANEL / 4; ++j4) {
        const float4 sc = *reinterpret_cast<const float4*>(colv + c0 + j4 * 4);
        const float4 sh = *reinterpret_cast<const float4*>(colv + COLV + c0 + j4 * 4);
        v[j4 * 4] = fmaf(v[j4 * 4], sc.x, sh.x); v[j4 * 4 + 1] = fmaf(v[j4 * 4 + 1], sc.y, sh.y);
        v[j4 * 4 + 2] = fmaf(v[j4 * 4 + 2], sc.z, sh.z); v[j4 * 4 + 3] = fmaf(v[j4 * 4 + 3], sc.w, sh.w);
      }
    }
    if (p.res_mode == 1) {
#pragma unroll
      for (int j = 0; j < PANEL; ++j) v[j] += rv[j];
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < PANEL; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (p.res_mode == 2) {
#pragma unroll
      for (int j = 0; j < PANEL; ++j) v[j] += rv[j];
    }
    if (p.red0 && live) {
      // cat = [red0 (red_c ch) | red1 (red_c ch)] ; out[col] += cat[2col] + cat[2col+1].  The source tensor is chosen per
      // 4-column group, so any red_c % 4 == 0 works (a 16-column panel may straddle red0 | red1: SCALING_RATIO 1 or 3).
      const int c2 = 2 * c0;
      const float* s0 = p.red0 + (size_t)r * p.ld_red0;
      const float* s1 = p.red1 + (size_t)r * p.ld_red1 - p.red_c;
#pragma unroll
      for (int j2 = 0; j2 < PANEL / 2; ++j2) {
        const int cc = c2 + j2 * 4;
        if (cc < 2 * p.red_c) {
          const float4 t = ldg_f4((cc < p.red_c ? s0 : s1) + cc);
          v[j2 * 2] += t.x + t.y;
          v[j2 * 2 + 1] += t.z + t.w;
        }
      }
    }
    if (w < PANEL) {
#pragma unroll
      for (int j = 0; j < PANEL; ++j)
        if (j >= w) v[j] = 0.f;
    }
  };
  auto emit_panel = [&](int c0, const float* v) {
    const int w = min(PANEL, p.cout - c0);
#pragma unroll
    for (int j4 = 0; j4 < PANEL / 4; ++j4)
      *reinterpret_cast<float4*>(my + j4 * 4) = masked ? make_float4(0.f, 0.f, 0.f, 0.f)
                                                       : make_float4(rnd(v[j4 * 4]), rnd(v[j4 * 4 + 1]), rnd(v[j4 * 4 + 2]), rnd(v[j4 * 4 + 3]));
    __syncwarp();
    panel_store(stg, p.out, p.ld_out, row0, rows_valid, c0, w, lane);
    __syncwarp();
  };

  if (p.n_ln == 0) {
    for (int c0 = 0; c0 < p.cout; c0 += PANEL) {
      float v[PANEL];
      finished_panel(c0, v);
      emit_panel(c0, v);
    }
    return;
  }
  // LayerNorm(s): exact two-pass statistics; finished values parked in TMEM over the accumulator columns
  float mean[2] = {0.f, 0.f}, rstd[2] = {1.f, 1.f};
  for (int ln = 0; ln < p.n_ln; ++ln) {
    float s1 = 0.f;
    for (int c0 = 0; c0 < p.cout; c0 += PANEL) {
      float v[PANEL];
      uint32_t raw[PANEL];
      if (ln == 0) {
        finished_panel(c0, v);
      } else {
        tmem_ld16(trow + c0, raw);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < PANEL; ++j)
          v[j] = (c0 + j < p.cout) ? (__uint_as_float(raw[j]) - mean[0]) * rstd[0] * colv[2 * COLV + c0 + j] + colv[3 * COLV + c0 + j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < PANEL; ++j) {
        s1 += v[j];
        raw[j] = __float_as_uint(v[j]);
      }
      tmem_st16(trow + c0, raw);
    }
    tmem_st_wait();
    const float m = s1 / (float)p.cout;
    float s2 = 0.f;
    for (int c0 = 0; c0 < p.cout; c0 += 16) {
      uint32_t raw[16];
      tmem_ld16(trow + c0, raw);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (c0 + j < p.cout) {
          const float d = __uint_as_float(raw[j]) - m;
          s2 = fmaf(d, d, s2);
        }
      }
    }
    mean[ln] = m;
    rstd[ln] = rsqrtf(s2 / (float)p.cout + p.ln_eps);
  }
  const int last = p.n_ln - 1;
  for (int c0 = 0; c0 < p.cout; c0 += PANEL) {
    uint32_t raw[PANEL];
    tmem_ld16(trow + c0, raw);
    tmem_ld_wait();
    float v[PANEL];
#pragma unroll
    for (int j = 0; j < PANEL; ++j)
      v[j] = (c0 + j < p.cout) ? (__uint_as_float(raw[j]) - mean[last]) * rstd[last] * colv[(2 + 2 * last) * COLV + c0 + j] +
                                     colv[(3 + 2 * last) * COLV + c0 + j]
                               : 0.f;
    emit_panel(c0, v);
  }
}

}  // namespace ls3d
