// Gather-GEMM with fused epilogues on tcgen05 tensor cores (sm_100a).
//
// One kernel family serves every dense contraction of the MSeg3D / SDSeg3D forward path:
//   * sparse 3-D convolution (SubM / strided / inverse) in output-stationary form
//       out[j,:] = epi( sum_k  in[nbr[k][j], :] . W[k] )          (replaces spconv's per-offset
//       gather -> cuBLAS mm -> scatter-add triple, reference call sites
//       det3d/models/backbones/scn_unet.py:15-20,39-46,89-160)
//   * every Linear(+BN)(+ReLU)(+residual)(+LayerNorm) of the point heads / TransVFE / SF-Phase
//       decoder (koff = 1, nbr = identity), reference det3d/models/point_heads/*.py
//   * the SF-Phase class-token cross attention as an epilogue of the q projection
//       (reference det3d/models/point_heads/context_module.py:320-376).
//
// Tile: 128 output rows x n_pad columns, fp32 accumulator in TMEM.  K loop over
// (kernel offset, 32-float channel chunk) "steps"; a step whose 128 rows have no neighbour is
// skipped.  Producer warps gather A rows (128-bit loads, cvt.rna.tf32) and the W[k] chunk into
// 128B-swizzled K-major shared tiles; one elected thread issues tcgen05.mma.kind::tf32; the
// producer warps then turn into the epilogue (tcgen05.ld -> affine/ReLU/residual/LayerNorm/
// attention -> global).
#include "common.cuh"
#include "../../include/ls3d.h"

namespace ls3d {

constexpr int TILE_M = 128;
constexpr int KCH = 32;         // floats per K chunk = one 128-byte swizzle row
constexpr int N_PROD = 128;     // producer / epilogue threads (4 warps)
constexpr int N_THREADS = 160;  // + 1 MMA warp
constexpr int MAX_KOFF = 27;
constexpr int MAX_TOK = 48;
constexpr int DHEAD = 24;

struct SmemLayout {
  uint32_t a_off[4], b_off[4];
  uint32_t nbr_off, flags_off, bar_off, tmem_slot_off, total;
};

__host__ __device__ inline uint32_t a_stage_bytes() { return TILE_M * 128; }
__host__ __device__ inline uint32_t b_stage_bytes(int n_pad) { return (uint32_t)n_pad * 128; }

__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t chunk) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4);
}

// SPLIT = true: error-compensated "3xTF32".  The tensor core truncates fp32 operands to tf32 (verified on B200), so with
//   x = x_hi + x_lo (x_hi = trunc_tf32(x), x_lo = x - x_hi exactly) and W = W_hi + W_lo (split on the host),
//   x.W ~= x_hi.W_hi + x_hi.W_lo + x_lo.W_hi   (dropped term x_lo.W_lo ~ 2^-22): fp32-level accuracy from three MMAs.
// x_hi comes for free (the raw fp32 tile, truncated by the MMA); the producers derive the x_lo tile from the landed raw tile.
template <int STAGES, bool SPLIT>
__global__ void __launch_bounds__(N_THREADS) gather_gemm_kernel(const ls3d_gemm_args p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A stages][B stages][nbr koff*128 ints][active flags][barriers][tmem slot]
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr uint32_t NSPLIT = SPLIT ? 2u : 1u;
  const uint32_t a_half = a_stage_bytes();            // one [128 x 32 float] tile
  const uint32_t b_half = b_stage_bytes(p.n_pad);     // one [n_pad x 32 float] tile
  const uint32_t a_bytes = NSPLIT * a_half;           // per stage: [A raw | A lo]
  const uint32_t b_bytes = NSPLIT * b_half;           // per stage: [W hi | W lo]
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + STAGES * a_bytes;
  int* nbr_s = (int*)(b_s + STAGES * b_bytes);
  uint32_t* act_s = (uint32_t*)(nbr_s + p.koff * TILE_M);  // [koff][4] warp ballots
  uint64_t* bars = (uint64_t*)(((uintptr_t)(act_s + p.koff * 4) + 7) & ~(uintptr_t)7);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int row0 = blockIdx.x * TILE_M;
  const int cin = p.c0 + p.c1;
  const int nchunk = (p.cin_pad + KCH - 1) / KCH;

  const uint32_t full_bar0 = smem_u32(bars);
  const uint32_t empty_bar0 = smem_u32(bars + STAGES);
  const uint32_t accum_bar = smem_u32(bars + 2 * STAGES);

  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)p.n_pad) tmem_cols <<= 1;

  // ---- setup
  if (tid < N_PROD) {
    const int r = row0 + tid;
    // all koff rulebook entries of this row are fetched before any is consumed (independent loads in flight)
    int jv[MAX_KOFF];
#pragma unroll
    for (int k = 0; k < MAX_KOFF; ++k) {
      jv[k] = -1;
      if (k < p.koff && r < p.m_out) jv[k] = p.nbr ? __ldg(p.nbr + (size_t)k * p.m_out + r) : r;
    }
#pragma unroll
    for (int k = 0; k < MAX_KOFF; ++k) {
      if (k < p.koff) {
        nbr_s[k * TILE_M + tid] = jv[k];
        uint32_t b = __ballot_sync(0xffffffffu, jv[k] >= 0);
        if (lane == 0) act_s[k * 4 + warp] = b;
      }
    }
  } else {
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(full_bar0 + 8 * s, N_PROD);
        mbar_init(empty_bar0 + 8 * s, 1);
      }
      mbar_init(accum_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  // total active steps (uniform across the CTA)
  int nsteps = 0;
  for (int k = 0; k < p.koff; ++k) {
    uint32_t any = act_s[k * 4] | act_s[k * 4 + 1] | act_s[k * 4 + 2] | act_s[k * 4 + 3];
    if (any) nsteps += nchunk;
  }

  if (tid < N_PROD) {
    // =========================== producer ===========================
    // cp.async (LDGSTS, zero-fill for missing rows) keeps up to STAGES-1 steps of gathers in flight per
    // thread; a step is published (fence.proxy.async + mbarrier arrive) once its own copies have landed.
    constexpr int DEPTH = STAGES - 1;
    int step = 0;
    const int ch = tid & 7;
    // make step j's tiles visible to the tensor core: (SPLIT) derive the x_lo tile from this thread's own landed
    // chunks of the raw tile, then cross-proxy fence + mbarrier arrive
    auto publish = [&](int j) {
      const int sj = j % STAGES;
      if (SPLIT) {
        uint8_t* a_raw = a_s + sj * a_bytes;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const uint32_t off = sw128(it * 16 + (tid >> 3), ch);
          float4 v = *reinterpret_cast<const float4*>(a_raw + off);
          v.x -= __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
          v.y -= __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
          v.z -= __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
          v.w -= __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
          *reinterpret_cast<float4*>(a_raw + a_half + off) = v;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(full_bar0 + 8 * sj);
    };
    for (int k = 0; k < p.koff; ++k) {
      uint32_t any = act_s[k * 4] | act_s[k * 4 + 1] | act_s[k * 4 + 2] | act_s[k * 4 + 3];
      if (!any) continue;
      const float* wk = p.w + (size_t)k * NSPLIT * p.n_pad * p.cin_pad;
      for (int c = 0; c < nchunk; ++c, ++step) {
        const int s = step % STAGES;
        const uint32_t ph = (uint32_t)(step / STAGES) & 1u;
        mbar_wait(empty_bar0 + 8 * s, ph ^ 1u);
        const uint32_t a_dst = smem_u32(a_s + s * a_bytes);
        const uint32_t b_dst = smem_u32(b_s + s * b_bytes);
        // ---- A: 128 rows x 8 chunks of 16 B; 8 lanes cover one row (4 full 128 B lines per warp request)
        const int col = c * KCH + ch * 4;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 16 + (tid >> 3);
          const int j = nbr_s[k * TILE_M + r];
          const bool ok = (j >= 0) && (col < cin);
          const float* src = p.in0;
          if (ok) src = (col < p.c0) ? (p.in0 + (size_t)j * p.ld0 + col) : (p.in1 + (size_t)j * p.ld1 + (col - p.c0));
          cp_async16(a_dst + sw128(r, ch), src, ok ? 16u : 0u);
        }
        // ---- B: n_pad rows x 8 chunks
        const int nb_it = p.n_pad / 16;
        for (int it = 0; it < nb_it; ++it) {
          const int idx = it * N_PROD + tid;
          const int n = idx >> 3;
          const int bcol = c * KCH + (idx & 7) * 4;
          const bool ok = bcol < p.cin_pad;
          cp_async16(b_dst + sw128(n, idx & 7), ok ? (wk + (size_t)n * p.cin_pad + bcol) : p.w, ok ? 16u : 0u);
          if (SPLIT)
            cp_async16(b_dst + b_half + sw128(n, idx & 7),
                       ok ? (wk + (size_t)(p.n_pad + n) * p.cin_pad + bcol) : p.w, ok ? 16u : 0u);
        }
        cp_async_commit();
        if (step >= DEPTH) {
          cp_async_wait<DEPTH>();
          publish(step - DEPTH);
        }
      }
    }
    // drain: publish the last min(DEPTH, nsteps) steps
    cp_async_wait<0>();
    for (int j = (nsteps > DEPTH ? nsteps - DEPTH : 0); j < nsteps; ++j) publish(j);
  } else {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc_tf32((uint32_t)p.n_pad);
    int step = 0;
    for (int k = 0; k < p.koff; ++k) {
      uint32_t any = act_s[k * 4] | act_s[k * 4 + 1] | act_s[k * 4 + 2] | act_s[k * 4 + 3];
      if (!any) continue;
      for (int c = 0; c < nchunk; ++c, ++step) {
        const int s = step % STAGES;
        const uint32_t ph = (uint32_t)(step / STAGES) & 1u;
        mbar_wait(full_bar0 + 8 * s, ph);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = make_desc_k_sw128(smem_u32(a_s + s * a_bytes));
          const uint64_t bdesc = make_desc_k_sw128(smem_u32(b_s + s * b_bytes));
          const uint64_t alo = make_desc_k_sw128(smem_u32(a_s + s * a_bytes + a_half));
          const uint64_t blo = make_desc_k_sw128(smem_u32(b_s + s * b_bytes + b_half));
          const int kc = min(KCH, p.cin_pad - c * KCH);  // multiple of 8
          for (int kk = 0; kk < kc / 8; ++kk) {
            // advance 8 tf32 = 32 B inside the 128 B swizzle row: +2 in the >>4 address field
            const uint64_t o = (uint64_t)(kk * 2);
            umma_tf32(tmem_acc, adesc + o, bdesc + o, idesc, (step > 0 || kk > 0) ? 1u : 0u);
            if (SPLIT) {
              umma_tf32(tmem_acc, adesc + o, blo + o, idesc, 1u);
              umma_tf32(tmem_acc, alo + o, bdesc + o, idesc, 1u);
            }
          }
          umma_commit(empty_bar0 + 8 * s);
          if (step == nsteps - 1) umma_commit(accum_bar);
        }
        __syncwarp();
      }
    }
  }

  // =========================== epilogue ===========================
  if (tid < N_PROD) {
    if (nsteps > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    const int r = row0 + tid;
    const bool live = r < p.m_out;
    const uint32_t trow = tmem_acc + ((uint32_t)(warp * 32) << 16);
    const bool have_acc = nsteps > 0;

    auto rnd = [&](float x) -> float { return p.round_out ? to_tf32(x) : x; };
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
        for (int d = 0; d < DHEAD; ++d) {
          float x = have_acc ? __uint_as_float(raw[d]) : 0.f;
          q[d] = x + (p.shift ? __ldg(p.shift + h * DHEAD + d) : 0.f);
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
        if (live) {
          float* dst = p.out + (size_t)r * p.ld_out + h * DHEAD;
#pragma unroll
          for (int d4 = 0; d4 < DHEAD / 4; ++d4)
            *reinterpret_cast<float4*>(dst + d4 * 4) =
                make_float4(rnd(o[d4 * 4] * inv), rnd(o[d4 * 4 + 1] * inv), rnd(o[d4 * 4 + 2] * inv), rnd(o[d4 * 4 + 3] * inv));
        }
      }
    } else {
      // value of column `col` after affine / residual / relu / channel-reduction
      auto finish = [&](float acc, int col) -> float {
        float x = acc;
        if (p.scale) x *= __ldg(p.scale + col);
        if (p.shift) x += __ldg(p.shift + col);
        if (p.res_mode == 1 && live) x += __ldg(p.res + (size_t)r * p.ld_res + col);
        if (p.relu) x = fmaxf(x, 0.f);
        if (p.res_mode == 2 && live) x += __ldg(p.res + (size_t)r * p.ld_res + col);
        if (p.red0 && live) {
          // cat = [red0 (red_c ch) | red1 (red_c ch)] ; out[col] += cat[2col] + cat[2col+1]
          const int c2 = 2 * col;
          const float* src = (c2 < p.red_c) ? (p.red0 + (size_t)r * p.ld_red0 + c2)
                                            : (p.red1 + (size_t)r * p.ld_red1 + (c2 - p.red_c));
          x += __ldg(src) + __ldg(src + 1);
        }
        return x;
      };
      const bool vec_ok = ((p.ld_out & 3) == 0) && ((p.cout & 3) == 0);
      const bool masked = p.row_mask && live && (__ldg(p.row_mask + (size_t)r * p.ld_mask) != 1.0f);
      float mean[2] = {0.f, 0.f}, rstd[2] = {1.f, 1.f};
      // LayerNorm statistics (up to two chained LayerNorms), exact two-pass form per LN
      for (int ln = 0; ln < p.n_ln; ++ln) {
        float s1 = 0.f;
        for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
          uint32_t raw[16];
          tmem_ld16(trow + c0, raw);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = c0 + j;
            if (col < p.cout) {
              float x = finish(have_acc ? __uint_as_float(raw[j]) : 0.f, col);
              if (ln == 1) x = (x - mean[0]) * rstd[0] * __ldg(p.ln_g0 + col) + __ldg(p.ln_b0 + col);
              s1 += x;
            }
          }
        }
        const float m = s1 / (float)p.cout;
        float s2 = 0.f;
        for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
          uint32_t raw[16];
          tmem_ld16(trow + c0, raw);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = c0 + j;
            if (col < p.cout) {
              float x = finish(have_acc ? __uint_as_float(raw[j]) : 0.f, col);
              if (ln == 1) x = (x - mean[0]) * rstd[0] * __ldg(p.ln_g0 + col) + __ldg(p.ln_b0 + col);
              const float d = x - m;
              s2 = fmaf(d, d, s2);
            }
          }
        }
        mean[ln] = m;
        rstd[ln] = rsqrtf(s2 / (float)p.cout + p.ln_eps);
      }
      for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
        uint32_t raw[16];
        tmem_ld16(trow + c0, raw);
        tmem_ld_wait();
        float y[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = c0 + j;
          float x = 0.f;
          if (col < p.cout) {
            x = finish(have_acc ? __uint_as_float(raw[j]) : 0.f, col);
            if (p.n_ln > 0) x = (x - mean[0]) * rstd[0] * __ldg(p.ln_g0 + col) + __ldg(p.ln_b0 + col);
            if (p.n_ln > 1) x = (x - mean[1]) * rstd[1] * __ldg(p.ln_g1 + col) + __ldg(p.ln_b1 + col);
          }
          y[j] = rnd(x);
        }
        if (masked) {
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] = 0.f;
        }
        if (live) {
          float* dst = p.out + (size_t)r * p.ld_out + c0;
          if (vec_ok) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
              if (c0 + j4 * 4 < p.cout)
                *reinterpret_cast<float4*>(dst + j4 * 4) =
                    make_float4(y[j4 * 4], y[j4 * 4 + 1], y[j4 * 4 + 2], y[j4 * 4 + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.cout) dst[j] = y[j];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_acc, tmem_cols);
}

static size_t smem_bytes_for(int stages, int n_pad, int koff, int nsplit) {
  size_t b = 1024;  // alignment slack
  b += (size_t)stages * nsplit * (a_stage_bytes() + b_stage_bytes(n_pad));
  b += (size_t)koff * TILE_M * 4 + (size_t)koff * 16;
  b += 8 + (2 * stages + 1) * 8 + 16;
  return b;
}

}  // namespace ls3d

extern "C" int ls3d_gather_gemm(const ls3d_gemm_args* a, void* stream) {
  using namespace ls3d;
  if (!a || !a->in0 || !a->w || !a->out) return LS3D_ERR_ARG;
  if (a->m_out <= 0) return LS3D_OK;
  if (a->koff < 1 || a->koff > MAX_KOFF) return LS3D_ERR_ARG;
  if (!a->nbr && a->koff != 1) return LS3D_ERR_ARG;
  if (a->n_pad % 16 || a->n_pad < 16 || a->n_pad > 256 || a->cout > a->n_pad) return LS3D_ERR_ARG;
  if (a->cin_pad % 8 || a->cin_pad < a->c0 + a->c1) return LS3D_ERR_ARG;
  if ((a->c0 & 3) || (a->c1 & 3) || (a->ld0 & 3) || (a->c1 && (a->ld1 & 3))) return LS3D_ERR_ARG;
  if (a->epi == LS3D_EPI_ATTN) {
    if (!a->attn_k || !a->attn_v || !a->frame_off || a->n_tok > MAX_TOK ||
        a->n_head * DHEAD != a->cout || (a->ld_out & 3))
      return LS3D_ERR_ARG;
  }
  if (a->n_ln < 0 || a->n_ln > 2) return LS3D_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ls3d_div_up(a->m_out, TILE_M);
  // deepest pipeline that still lets two CTAs share one SM (<= ~110 KB each); tiles too big for that get one CTA per
  // SM and as many stages as fit
  const int nsplit = a->precise ? 2 : 1;
  int stages = 4;
  while (stages > 2 && smem_bytes_for(stages, a->n_pad, a->koff, nsplit) > 110 * 1024) --stages;
  if (smem_bytes_for(stages, a->n_pad, a->koff, nsplit) > 110 * 1024) {
    stages = 4;
    while (stages > 2 && smem_bytes_for(stages, a->n_pad, a->koff, nsplit) > 220 * 1024) --stages;
  }
  const size_t smem = smem_bytes_for(stages, a->n_pad, a->koff, nsplit);
  if (smem > 227 * 1024) return LS3D_ERR_ARG;
  cudaError_t e;
#define LS3D_GG_LAUNCH(S, P)                                                                         \
  {                                                                                                  \
    e = cudaFuncSetAttribute(gather_gemm_kernel<S, P>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                             (int)smem);                                                             \
    if (e != cudaSuccess) return (int)e;                                                             \
    gather_gemm_kernel<S, P><<<grid, N_THREADS, smem, st>>>(*a);                                     \
  }
  if (a->precise) {
    if (stages == 4) LS3D_GG_LAUNCH(4, true)
    else if (stages == 3) LS3D_GG_LAUNCH(3, true)
    else LS3D_GG_LAUNCH(2, true)
  } else {
    if (stages == 4) LS3D_GG_LAUNCH(4, false)
    else if (stages == 3) LS3D_GG_LAUNCH(3, false)
    else LS3D_GG_LAUNCH(2, false)
  }
#undef LS3D_GG_LAUNCH
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}
