// SF-Phase point decoder: all TransformerDecoderLayer.forward_post steps of the POINT stream in ONE persistent launch.
//
// Replaces, per decoder layer, the point side of reference det3d/models/point_heads/context_module.py:
//   211-250 (forward_post: tgt = norm2(tgt + crossocr_attn(tgt, memory)); tgt = norm3(tgt + linear2(relu(linear1(tgt)))))
//   320-376 (SparsePointCorssAttention: q_proj, per-frame softmax(q k^T / sqrt(dh)) v over the class tokens, out_proj)
//   147-171 (TransformerDecoder: the layer loop and the closing norm_tgt)
// which the unfused path ran as 5 launches per layer (q GEMM, attention, out GEMM + LN, FFN GEMM, FFN GEMM + LN) with a
// full HBM round trip of the [points x 96] activations between each.  Here a CTA owns 128-point tiles; the activations of
// a tile never leave the SM between the first load and the last store:
//   * tgt (fp32) lives in TENSOR MEMORY columns [384, 480); every GEMM's A operand is the bf16 hi / lo split of the
//     current activations, written by the row threads into TMEM columns [192, 384) (tcgen05.st) and read from there by the
//     MMA (TS mode); accumulators are TMEM columns [0, 192);
//   * the weights of the four Linears of a layer (error-compensated bf16x3 images, the gather-GEMM's PackedWeight chunks:
//     221 KB per layer) stream from L2 through a 5-slot shared-memory ring (cp.async.bulk + mbarriers), in consumption
//     order, by a loader thread that runs ahead of the MMA issuer;
//   * the class-token K / V of the tile's frame and layer (2 x 13 KB) are double-buffered in shared memory; rows of another
//     frame (a tile straddling a frame boundary) read theirs from global memory;
//   * 16 row warps: warp w serves the 32 points of TMEM lane quarter w & 3 and column group (= attention head) w >> 2,
//     i.e. 24 of the 96 channels (48 of the 192 FFN channels) of each of its points: bias, the 34-token cross attention of
//     its head on the CUDA cores (fp32), residual, exact two-pass LayerNorm (row sums exchanged through shared memory between
//     the four warps of a point), ReLU, operand split; 1 MMA-issue warp; 1 loader warp.  (Four warps per scheduler: a single
//     row warp per scheduler left every shared-memory / tensor-memory / FMA latency exposed - 1.32 ms per launch against 0.83.)
// Measured bound (ncu, profiles/r02_ncu_sffm_decoder*): the attention phase - one broadcast LDS.128 of K / V per 4 FMAs keeps
// the shared-memory pipe, not the FMA pipe, busy; the four GEMM phases of a layer cost ~5.2 k cycles of tensor pipe per tile.
// Arithmetic is the unfused path's: x_hi.W_hi + x_hi.W_lo + x_lo.W_hi bf16 products with fp32 accumulation (~2^-17 per
// product), fp32 softmax / LayerNorm.
#include "common.cuh"
#include "../../include/ls3d.h"

namespace ls3d {
namespace dec {

constexpr int E = 96, FF = 192, DH = 24, NH = 4, TILE = 128;
constexpr int ROW_WARPS = 16, MMA_WARP = 16, LOAD_WARP = 17, N_THREADS = 18 * 32;   // row warp w: TMEM lane quarter w & 3, column group w >> 2
constexpr int RING = 5, SLOT_BYTES = 24576;
constexpr int CH_E = E * 2 * 64;        // stacked chunk: [W_hi ; W_lo] rows of 64 bytes (32 bf16), SWIZZLE_64B
constexpr int CH_F = FF * 128;          // wide chunk: rows [hi 32 | lo 32] of 128 bytes, SWIZZLE_128B
constexpr int CHUNKS_PER_LAYER = 3 + 3 + 3 + 6;
constexpr int LAYER_W_BYTES = 3 * CH_E + 3 * CH_E + 3 * CH_F + 6 * CH_E;
// per-layer vector block (floats): q bias, out-proj bias, norm2 gamma / beta, linear1 bias, linear2 bias, norm3 gamma / beta
constexpr int V_BQ = 0, V_BO = 96, V_G2 = 192, V_B2 = 288, V_B1 = 384, V_BF2 = 576, V_G3 = 672, V_B3 = 768, VEC_LAYER = 864;
constexpr int COL_ACC = 0, COL_A = 192, COL_TGT = 384;
constexpr int MAX_LAYERS = 8, MAX_TOK = 64;

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3])
               : "memory");
}

struct Args {
  const float* in;
  int ld_in, n;
  const uint8_t* w;          // [n_layer][LAYER_W_BYTES]
  const float* vec;          // [n_layer][VEC_LAYER] + final norm gamma[E], beta[E]
  const float* k;            // [n_layer][n_frames][NH][n_tok][DH]
  const float* v;
  const int* frame_off;      // [n_frames] first row of each frame
  int n_frames, n_tok, n_layer, final_norm;
  float scale, eps;
  float* out;
  int ld_out;
};

__device__ __forceinline__ int frame_of(const int* off, int nf, int r) {
  int f = 0;
  for (int i = 1; i < nf; ++i)
    if (r >= __ldg(off + i)) f = i;
  return f;
}

// 16 fp32 -> 8 packed bf16x2 hi words + 8 lo words (x = hi + lo to ~2^-17)
__device__ __forceinline__ void split16(const float* x, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t h = pack_bf16x2(x[2 * j], x[2 * j + 1]);
    hi[j] = h;
    lo[j] = pack_bf16x2(x[2 * j] - __uint_as_float(h << 16), x[2 * j + 1] - __uint_as_float(h & 0xFFFF0000u));
  }
}

// softmax(q . K^T * scale) V over L tokens of one head; K / V rows of DH floats (shared or global memory)
__device__ __forceinline__ void attend(const float* q, const float* kh, const float* vh, int L, float scale, float* o) {
#pragma unroll
  for (int d = 0; d < DH; ++d) o[d] = 0.f;
  float mx = -INFINITY, den = 0.f;
  for (int l0 = 0; l0 < L; l0 += 8) {
    float s[8];
    float cm = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = -INFINITY;
      if (l0 + j < L) {
        const float4* kr = reinterpret_cast<const float4*>(kh + (size_t)(l0 + j) * DH);
        a = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < DH / 4; ++d4) {
          const float4 kv = kr[d4];
          a = fmaf(q[d4 * 4 + 0], kv.x, a);
          a = fmaf(q[d4 * 4 + 1], kv.y, a);
          a = fmaf(q[d4 * 4 + 2], kv.z, a);
          a = fmaf(q[d4 * 4 + 3], kv.w, a);
        }
        a *= scale;
      }
      s[j] = a;
      cm = fmaxf(cm, a);
    }
    const float mn = fmaxf(mx, cm);
    const float corr = __expf(mx - mn);
    den *= corr;
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] *= corr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (l0 + j < L) {
        const float e = __expf(s[j] - mn);
        den += e;
        const float4* vr = reinterpret_cast<const float4*>(vh + (size_t)(l0 + j) * DH);
#pragma unroll
        for (int d4 = 0; d4 < DH / 4; ++d4) {
          const float4 vv = vr[d4];
          o[d4 * 4 + 0] = fmaf(e, vv.x, o[d4 * 4 + 0]);
          o[d4 * 4 + 1] = fmaf(e, vv.y, o[d4 * 4 + 1]);
          o[d4 * 4 + 2] = fmaf(e, vv.z, o[d4 * 4 + 2]);
          o[d4 * 4 + 3] = fmaf(e, vv.w, o[d4 * 4 + 3]);
        }
      }
    }
    mx = mn;
  }
  const float inv = 1.f / den;
#pragma unroll
  for (int d = 0; d < DH; ++d) o[d] *= inv;
}

__global__ void __launch_bounds__(N_THREADS, 1) sffm_decoder_kernel(const Args p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int kv_floats = NH * p.n_tok * DH;                     // floats of K (and of V) per (layer, frame)
  uint8_t* ring_s = smem;                                       // [RING][SLOT_BYTES]
  float* kv_s = reinterpret_cast<float*>(ring_s + RING * SLOT_BYTES);     // [2][K | V]
  float* vec_s = kv_s + 2 * 2 * kv_floats;                      // [n_layer][VEC_LAYER] + [2][E]
  float* red_s = vec_s + p.n_layer * VEC_LAYER + 2 * E;        // [2][4][TILE] LayerNorm partial sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(((uintptr_t)(red_s + 2 * 4 * TILE) + 7) & ~(uintptr_t)7);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * RING + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t wfull0 = smem_u32(bars);                      // weight chunk landed   [RING]
  const uint32_t wempty0 = smem_u32(bars + RING);              // weight slot consumed  [RING] (tcgen05.commit)
  const uint32_t aready = smem_u32(bars + 2 * RING);           // A operand written (4 row warps)
  const uint32_t accfull = smem_u32(bars + 2 * RING + 1);      // accumulator complete (tcgen05.commit)
  const uint32_t kvfull0 = smem_u32(bars + 2 * RING + 2);      // K / V landed   [2]
  const uint32_t kvempty0 = smem_u32(bars + 2 * RING + 4);     // K / V consumed [2] (4 row warps)
  const int ntiles = (p.n + TILE - 1) / TILE;

  for (int i = tid; i < p.n_layer * VEC_LAYER + 2 * E; i += N_THREADS) vec_s[i] = __ldg(p.vec + i);
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < RING; ++s) {
        mbar_init(wfull0 + 8 * s, 1);
        mbar_init(wempty0 + 8 * s, 1);
      }
      mbar_init(aready, ROW_WARPS);
      mbar_init(accfull, 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(kvfull0 + 8 * b, 1);
        mbar_init(kvempty0 + 8 * b, ROW_WARPS);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < ROW_WARPS) {
    // =========================== row warps ===========================
    const int q = warp & 3, h = warp >> 2;                      // TMEM lane quarter; column group = attention head
    const int row = q * 32 + lane;
    const int cg = h * DH;                                      // first of this thread's 24 channels
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t_acc = tl + COL_ACC, t_a = tl + COL_A, t_tgt = tl + COL_TGT;
    float* red_a = red_s + h * TILE + row;                      // [4][TILE]: this thread's slot; the row's four are TILE apart
    float* red_b = red_a + 4 * TILE;
    uint32_t acc_ph = 0;
    uint32_t g = 0;                                             // (tile, layer) counter: K / V buffer g & 1
    auto bar_rows = [] { asm volatile("bar.sync 1, 512;" ::: "memory"); };

    // 24 channels -> the A operand slots (bf16 hi / lo): three 8-channel groups; group g8 of the row -> chunk g8 / 4,
    // 4 TMEM columns at 4 (g8 % 4) (+ 16 for the lo half)
    auto store_a24 = [&](const float* x) {
#pragma unroll
      for (int gq = 0; gq < 3; ++gq) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float x0 = x[8 * gq + 2 * j], x1 = x[8 * gq + 2 * j + 1];
          const uint32_t hh = pack_bf16x2(x0, x1);
          hi[j] = hh;
          lo[j] = pack_bf16x2(x0 - __uint_as_float(hh << 16), x1 - __uint_as_float(hh & 0xFFFF0000u));
        }
        const int g8 = 3 * h + gq;
        const uint32_t col = (uint32_t)(32 * (g8 >> 2) + 4 * (g8 & 3));
        tmem_st4(t_a + col, hi);
        tmem_st4(t_a + col + 16, lo);
      }
    };
    auto publish = [&] {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(aready);
    };
    // tgt -> TMEM (fp32) and the A operand slots; then publish
    auto publish_tgt = [&](const float* v) {
#pragma unroll
      for (int gq = 0; gq < 3; ++gq) {
        uint32_t raw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) raw[j] = __float_as_uint(v[8 * gq + j]);
        tmem_st8(t_tgt + cg + 8 * gq, raw);
      }
      store_a24(v);
      publish();
    };
    // v = acc (two stacked halves) + bias + tgt for this thread's 24 channels
    auto residual_row = [&](float* v, const float* bias) {
      uint32_t a[24], b[24], t[24];
#pragma unroll
      for (int gq = 0; gq < 3; ++gq) {
        tmem_ld8(t_acc + cg + 8 * gq, a + 8 * gq);
        tmem_ld8(t_acc + E + cg + 8 * gq, b + 8 * gq);
        tmem_ld8(t_tgt + cg + 8 * gq, t + 8 * gq);
      }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < DH; ++j)
        v[j] = (__uint_as_float(a[j]) + __uint_as_float(b[j])) + bias[cg + j] + __uint_as_float(t[j]);
    };
    // exact two-pass LayerNorm over the row's 96 channels held by four threads (fixed summation order)
    auto layer_norm = [&](float* v, const float* gm, const float* bt) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < DH; ++j) s += v[j];
      *red_a = s;
      bar_rows();
      const float* ra = red_s + row;
      const float m = ((ra[0] + ra[TILE]) + (ra[2 * TILE] + ra[3 * TILE])) / (float)E;
      float s2 = 0.f;
#pragma unroll
      for (int j = 0; j < DH; ++j) {
        const float d = v[j] - m;
        s2 = fmaf(d, d, s2);
      }
      *red_b = s2;
      bar_rows();
      const float* rb = ra + 4 * TILE;
      const float rstd = rsqrtf(((rb[0] + rb[TILE]) + (rb[2 * TILE] + rb[3 * TILE])) / (float)E + p.eps);
#pragma unroll
      for (int j = 0; j < DH; ++j) v[j] = (v[j] - m) * rstd * gm[cg + j] + bt[cg + j];
    };

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int r = tile * TILE + row;
      const bool live = r < p.n;
      const int f0 = frame_of(p.frame_off, p.n_frames, tile * TILE);
      const int f = live ? frame_of(p.frame_off, p.n_frames, r) : f0;
      float v[DH];
      {
        const float* src = p.in + (size_t)r * p.ld_in + cg;
#pragma unroll
        for (int c = 0; c < DH / 4; ++c) {
          const float4 t = live ? ldg_f4(src + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
        }
      }
      publish_tgt(v);
      for (int l = 0; l < p.n_layer; ++l, ++g) {
        const float* vl = vec_s + l * VEC_LAYER;
        // ---------------- q projection -> cross attention of head h over the class tokens -> A operand
        mbar_wait(accfull, acc_ph);
        acc_ph ^= 1u;
        tc_fence_after();
        const int kb = g & 1;
        mbar_wait(kvfull0 + 8 * kb, (g >> 1) & 1u);
        const float* ks = kv_s + kb * 2 * kv_floats;
        const float* vs = ks + kv_floats;
        {
          uint32_t a[24], b[24];
#pragma unroll
          for (int gq = 0; gq < 3; ++gq) {
            tmem_ld8(t_acc + cg + 8 * gq, a + 8 * gq);
            tmem_ld8(t_acc + E + cg + 8 * gq, b + 8 * gq);
          }
          tmem_ld_wait();
          float qv[DH], o[DH];
#pragma unroll
          for (int d = 0; d < DH; ++d) qv[d] = (__uint_as_float(a[d]) + __uint_as_float(b[d])) + vl[V_BQ + cg + d];
          if (f == f0) {
            attend(qv, ks + h * p.n_tok * DH, vs + h * p.n_tok * DH, p.n_tok, p.scale, o);
          } else {
            const size_t goff = ((size_t)l * p.n_frames + f) * kv_floats + (size_t)h * p.n_tok * DH;
            attend(qv, p.k + goff, p.v + goff, p.n_tok, p.scale, o);
          }
          store_a24(o);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(kvempty0 + 8 * kb);
        publish();
        // ---------------- out projection + residual + norm2
        mbar_wait(accfull, acc_ph);
        acc_ph ^= 1u;
        tc_fence_after();
        residual_row(v, vl + V_BO);
        layer_norm(v, vl + V_G2, vl + V_B2);
        publish_tgt(v);
        // ---------------- linear1 + ReLU -> A operand (this warp: FFN channels [48 h, 48 h + 48) = panels 3 h .. 3 h + 2)
        mbar_wait(accfull, acc_ph);
        acc_ph ^= 1u;
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int pn = 3 * h + i;
          uint32_t a[16], hi[8], lo[8];
          tmem_ld16(t_acc + 16 * pn, a);
          tmem_ld_wait();
          float x[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = fmaxf(__uint_as_float(a[j]) + vl[V_B1 + 16 * pn + j], 0.f);
          split16(x, hi, lo);
          const uint32_t col = (uint32_t)(32 * (pn >> 1) + 8 * (pn & 1));
          tmem_st8(t_a + col, hi);
          tmem_st8(t_a + col + 16, lo);
        }
        publish();
        // ---------------- linear2 + residual + norm3
        mbar_wait(accfull, acc_ph);
        acc_ph ^= 1u;
        tc_fence_after();
        residual_row(v, vl + V_BF2);
        layer_norm(v, vl + V_G3, vl + V_B3);
        if (l + 1 < p.n_layer) publish_tgt(v);
      }
      if (p.final_norm) layer_norm(v, vec_s + p.n_layer * VEC_LAYER, vec_s + p.n_layer * VEC_LAYER + E);
      if (live) {
        float* dst = p.out + (size_t)r * p.ld_out + cg;
#pragma unroll
        for (int c = 0; c < DH / 4; ++c)
          *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc_e = make_idesc_bf16((uint32_t)E);
    const uint32_t idesc_2e = make_idesc_bf16(2u * (uint32_t)E);
    const uint32_t tbase = bcast0(tmem_base);
    const uint32_t tacc = tbase + COL_ACC, ta0 = tbase + COL_A;
    const uint32_t ring0 = smem_u32(ring_s);
    uint32_t ar_ph = 0, w_ph = 0;
    int slot = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int l = 0; l < p.n_layer; ++l) {
#pragma unroll 1
        for (int phase = 0; phase < 4; ++phase) {               // q, out, linear1, linear2
          const int nchunk = phase == 3 ? 6 : 3;
          const bool wide = phase == 2;
          mbar_wait(aready, ar_ph);
          ar_ph ^= 1u;
          tc_fence_after();
          for (int c = 0; c < nchunk; ++c) {
            mbar_wait(wfull0 + 8 * slot, w_ph);
            tc_fence_after();
            const uint32_t ws = ring0 + (uint32_t)slot * SLOT_BYTES;
            const uint32_t a_hi = ta0 + (uint32_t)(32 * c), a_lo = a_hi + 16;
            if (elect_one()) {
              if (wide) {
                const uint64_t bdesc = make_desc_k_sw128(ws);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const uint64_t o = (uint64_t)(2 * j);
                  umma_bf16_ts(tacc, a_hi + 8 * j, bdesc + o, idesc_2e, (c > 0 || j > 0) ? 1u : 0u);
                  umma_bf16_ts(tacc, a_hi + 8 * j, bdesc + 4 + o, idesc_2e, 1u);
                  umma_bf16_ts(tacc, a_lo + 8 * j, bdesc + o, idesc_2e, 1u);
                }
              } else {
                const uint64_t bdesc = make_desc_k_sw64(ws);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const uint64_t o = (uint64_t)(2 * j);
                  umma_bf16_ts(tacc, a_hi + 8 * j, bdesc + o, idesc_2e, (c > 0 || j > 0) ? 1u : 0u);
                  umma_bf16_ts(tacc, a_lo + 8 * j, bdesc + o, idesc_e, 1u);
                }
              }
              umma_commit(wempty0 + 8 * slot);
              if (c == nchunk - 1) umma_commit(accfull);
            }
            __syncwarp();
            if (++slot == RING) {
              slot = 0;
              w_ph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == LOAD_WARP) {
    // =========================== weight / K / V loader ===========================
    if (lane == 0) {
      uint32_t w_ph = 0, g = 0;
      int slot = 0;
      const uint32_t ring0 = smem_u32(ring_s);
      const uint32_t kv_bytes = (uint32_t)kv_floats * 4u;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int f0 = frame_of(p.frame_off, p.n_frames, tile * TILE);
        for (int l = 0; l < p.n_layer; ++l, ++g) {
          const int kb = g & 1;
          mbar_wait(kvempty0 + 8 * kb, ((g >> 1) & 1u) ^ 1u);
          mbar_arrive_expect_tx(kvfull0 + 8 * kb, 2u * kv_bytes);
          const size_t goff = ((size_t)l * p.n_frames + f0) * kv_floats;
          bulk_g2s(smem_u32(kv_s + kb * 2 * kv_floats), p.k + goff, kv_bytes, kvfull0 + 8 * kb);
          bulk_g2s(smem_u32(kv_s + kb * 2 * kv_floats + kv_floats), p.v + goff, kv_bytes, kvfull0 + 8 * kb);
          const uint8_t* src = p.w + (size_t)l * LAYER_W_BYTES;
          for (int c = 0; c < CHUNKS_PER_LAYER; ++c) {
            const uint32_t bytes = (c >= 6 && c < 9) ? (uint32_t)CH_F : (uint32_t)CH_E;
            mbar_wait(wempty0 + 8 * slot, w_ph ^ 1u);
            mbar_arrive_expect_tx(wfull0 + 8 * slot, bytes);
            bulk_g2s(ring0 + (uint32_t)slot * SLOT_BYTES, src, bytes, wfull0 + 8 * slot);
            src += bytes;
            if (++slot == RING) {
              slot = 0;
              w_ph ^= 1u;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, 512);
}

static size_t smem_bytes_for(int n_layer, int n_tok) {
  size_t b = 1024 + (size_t)RING * SLOT_BYTES;
  b += (size_t)2 * 2 * NH * n_tok * DH * 4;
  b += (size_t)(n_layer * VEC_LAYER + 2 * E + 2 * 4 * TILE) * 4 + 8;
  b += (2 * RING + 6) * 8 + 16;
  return b;
}

}  // namespace dec
}  // namespace ls3d

extern "C" int ls3d_sffm_decoder_weight_bytes(int32_t n_layer, int64_t* w_bytes, int64_t* vec_floats) {
  using namespace ls3d::dec;
  if (n_layer < 1 || n_layer > MAX_LAYERS || !w_bytes || !vec_floats) return LS3D_ERR_ARG;
  *w_bytes = (int64_t)n_layer * LAYER_W_BYTES;
  *vec_floats = (int64_t)n_layer * VEC_LAYER + 2 * E;
  return LS3D_OK;
}

extern "C" int ls3d_sffm_decoder(const float* tgt_in, int32_t ld_in, int32_t n, const void* w, const float* vec, const float* k,
                                 const float* v, const int32_t* frame_off, int32_t n_frames, int32_t n_tok, int32_t n_layer,
                                 int32_t n_head, int32_t d_model, int32_t d_ffn, int32_t final_norm, float attn_scale,
                                 float ln_eps, float* out, int32_t ld_out, void* stream) {
  using namespace ls3d;
  using namespace ls3d::dec;
  if (n <= 0) return LS3D_OK;
  if (!tgt_in || !w || !vec || !k || !v || !frame_off || !out) return LS3D_ERR_ARG;
  if (d_model != E || d_ffn != FF || n_head != NH || n_layer < 1 || n_layer > MAX_LAYERS || n_tok < 1 || n_tok > MAX_TOK ||
      n_frames < 1)
    return LS3D_ERR_ARG;
  if ((ld_in & 3) || (ld_out & 3) || ((uintptr_t)tgt_in & 15) || ((uintptr_t)out & 15) || ((uintptr_t)w & 15) ||
      ((uintptr_t)k & 15) || ((uintptr_t)v & 15) || ((n_tok * DH * NH * 4) & 15))
    return LS3D_ERR_ARG;
  const size_t smem = smem_bytes_for(n_layer, n_tok);
  if (smem > 227 * 1024) return LS3D_ERR_ARG;
  static bool optin[64] = {false};
  cudaError_t e = ls3d_optin_smem(sffm_decoder_kernel, optin);
  if (e != cudaSuccess) return (int)e;
  Args a;
  a.in = tgt_in; a.ld_in = ld_in; a.n = n; a.w = (const uint8_t*)w; a.vec = vec; a.k = k; a.v = v; a.frame_off = frame_off;
  a.n_frames = n_frames; a.n_tok = n_tok; a.n_layer = n_layer; a.final_norm = final_norm; a.scale = attn_scale; a.eps = ln_eps;
  a.out = out; a.ld_out = ld_out;
  const int ntiles = ls3d_div_up(n, TILE);
  const int num_sms = ls3d_num_sms();
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  sffm_decoder_kernel<<<grid, N_THREADS, smem, (cudaStream_t)stream>>>(a);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}
