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
// Persistent, warp-specialised CTA (one per SM), 128-row output tiles taken round-robin:
//   warps 0-7   producers : per (kernel offset, 32-float channel chunk) "step" gather the tile's input rows and the W[k]
//                           chunk with cp.async (zero fill for missing rows) into 128B-swizzled K-major smem stages and
//                           hand completion to the "landed" mbarrier with cp.async.mbarrier.arrive.noinc - they never
//                           wait for data, so all stages stay in flight and the ring runs ahead across tile boundaries;
//                           steps whose 128 rows have no neighbour are skipped
//   warps 8-11  splitters : wait "landed", (3xTF32) derive the x_lo tile from the raw tile, cross-proxy fence, publish "full"
//   warp  12    MMA issuer: one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=n_pad, K=8) into one of two TMEM
//                           accumulator buffers; tcgen05.commit releases smem stages / publishes the accumulator
//   warps 13-16 epilogue  : tcgen05.ld -> folded BN / ReLU / residual / channel reduction / row mask / LayerNorm(s) /
//                           class-token attention -> shared-memory panel -> coalesced global stores; overlaps the next
//                           tile's main loop (double-buffered TMEM)
#include "gemm_epilogue.cuh"

namespace ls3d {

constexpr int N_PROD_WARPS = 8;
constexpr int N_PROD = N_PROD_WARPS * 32;
constexpr int SPLIT_WARP0 = N_PROD_WARPS;     // 4 splitter warps = one thread per tile row
constexpr int MMA_WARP = N_PROD_WARPS + 4;
constexpr int N_THREADS = (N_PROD_WARPS + 4 + 1 + 4) * 32;

__host__ __device__ inline uint32_t a_stage_bytes() { return TILE_M * 128; }
__host__ __device__ inline uint32_t b_stage_bytes(int n_pad) { return (uint32_t)n_pad * 128; }

__device__ __forceinline__ void bar_sync_producers() { asm volatile("bar.sync 1, %0;" ::"n"(N_PROD) : "memory"); }

// SPLIT = true: error-compensated "3xTF32".  The tensor core truncates fp32 operands to tf32 (verified on B200), so with
//   x = x_hi + x_lo (x_hi = trunc_tf32(x), x_lo = x - x_hi exactly) and W = W_hi + W_lo (split on the host),
//   x.W ~= x_hi.W_hi + x_hi.W_lo + x_lo.W_hi   (dropped term x_lo.W_lo ~ 2^-22): fp32-level accuracy from three MMAs.
// x_hi comes for free (the raw fp32 tile, truncated by the MMA); the producers derive the x_lo tile from the landed raw tile.
template <int STAGES, bool SPLIT>
__global__ void __launch_bounds__(N_THREADS, 1) gather_gemm_kernel(const ls3d_gemm_args p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr uint32_t NSPLIT = SPLIT ? 2u : 1u;
  const uint32_t a_half = a_stage_bytes();            // one [128 x 32 float] tile
  const uint32_t b_half = b_stage_bytes(p.n_pad);     // one [n_pad x 32 float] tile
  const uint32_t a_bytes = NSPLIT * a_half;           // per stage: [A raw | A lo]
  const uint32_t b_bytes = NSPLIT * b_half;           // per stage: [W hi | W lo]
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + STAGES * a_bytes;
  int* nbr_s = (int*)(b_s + STAGES * b_bytes);                     // [2][koff][128]
  uint32_t* act_s = (uint32_t*)(nbr_s + 2 * p.koff * TILE_M);      // [2][koff][4] warp ballots
  uint32_t* mask_s = act_s + 2 * p.koff * 4;                       // [MASK_RING] active-offset masks
  uint64_t* bars = (uint64_t*)(((uintptr_t)(mask_s + MASK_RING) + 7) & ~(uintptr_t)7);
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * STAGES + 4 + MASK_RING);
  float* colv = (float*)(((uintptr_t)(tmem_slot + 4) + 15) & ~(uintptr_t)15);      // [6][COLV] per-column vectors
  float* stg = colv + 6 * COLV;                                                      // [128][STG_LD] epilogue panel

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cin = p.c0 + p.c1;
  const int nchunk = (p.cin_pad + KCH - 1) / KCH;
  const int ntiles = (p.m_out + TILE_M - 1) / TILE_M;

  const uint32_t full_bar0 = smem_u32(bars);
  const uint32_t empty_bar0 = smem_u32(bars + STAGES);
  const uint32_t land_bar0 = smem_u32(bars + 2 * STAGES);       // cp.async data landed [STAGES]
  const uint32_t accf_bar0 = smem_u32(bars + 3 * STAGES);       // accumulator full  [2]
  const uint32_t acce_bar0 = smem_u32(bars + 3 * STAGES + 2);   // accumulator empty [2]
  const uint32_t mask_bar0 = smem_u32(bars + 3 * STAGES + 4);   // mask published    [MASK_RING]

  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * (uint32_t)p.n_pad) tmem_cols <<= 1;

  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(full_bar0 + 8 * s, 128);
        mbar_init(empty_bar0 + 8 * s, 1);
        mbar_init(land_bar0 + 8 * s, N_PROD);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(accf_bar0 + 8 * b, 1);
        mbar_init(acce_bar0 + 8 * b, 128);
      }
      for (int m = 0; m < MASK_RING; ++m) mbar_init(mask_bar0 + 8 * m, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool poll = (p.debug_skip & 1024) != 0;
  auto WAIT = [&](uint32_t bar, uint32_t parity) {
    if (poll) mbar_wait_poll(bar, parity); else mbar_wait(bar, parity);
  };

  if (warp < N_PROD_WARPS) {
    // =========================== producers ===========================
    const int ch = tid & 7;
    const int rsub = tid >> 3;                         // row within a 32-row group (A) / 32-row group of W
    constexpr int A_IT = TILE_M * 8 / N_PROD;
    constexpr int B_IT = 256 * 8 / N_PROD;
    uint32_t soff[B_IT];                               // swizzled smem offsets of (row it*32 + rsub, chunk ch)
#pragma unroll
    for (int it = 0; it < B_IT; ++it) soff[it] = sw128(it * (N_PROD / 8) + rsub, ch);
    const int nbg = (p.n_pad + 31) / 32;               // 32-row groups of the W tile
    uint32_t bmask = 0;                                // groups in which this thread's row exists
    for (int it = 0; it < nbg; ++it)
      if (it * 32 + rsub < p.n_pad) bmask |= 1u << it;
    const size_t wstride = (size_t)32 * p.cin_pad;     // floats between row groups
    const size_t wsplit = (size_t)p.n_pad * p.cin_pad; // floats from W_hi to W_lo
    int g = 0;                                        // global step counter (ring position), continues across tiles
    int ti = 0;
    const int my_row = tid & (TILE_M - 1);            // thread t loads rulebook row (t & 127) for half of the offsets
    const int khalf = (p.koff + 1) / 2;
    const int k_lo = (tid < TILE_M) ? 0 : khalf;
    const int k_hi = (tid < TILE_M) ? khalf : p.koff;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1;
      const int r = tile * TILE_M + my_row;
      // all rulebook entries are fetched before any is consumed (independent loads in flight)
      int jv[(MAX_KOFF + 1) / 2];
#pragma unroll
      for (int q = 0; q < (MAX_KOFF + 1) / 2; ++q) {
        const int k = k_lo + q;
        jv[q] = -1;
        if (k < k_hi && r < p.m_out) jv[q] = p.nbr ? __ldg(p.nbr + (size_t)k * p.m_out + r) : r;
      }
      bar_sync_producers();                           // everyone finished issuing the tile that used nbr_s[buf] before
      int* nb = nbr_s + buf * p.koff * TILE_M;
      uint32_t* ac = act_s + buf * p.koff * 4;
#pragma unroll
      for (int q = 0; q < (MAX_KOFF + 1) / 2; ++q) {
        const int k = k_lo + q;
        if (k < k_hi) {
          nb[k * TILE_M + (my_row & 31) * 4 + (my_row >> 5)] = jv[q];
          const uint32_t b = __ballot_sync(0xffffffffu, jv[q] >= 0);
          if (lane == 0) ac[k * 4 + (warp & 3)] = b;
        }
      }
      bar_sync_producers();
      uint32_t mask = 0;
      for (int k = 0; k < p.koff; ++k)
        if (ac[k * 4] | ac[k * 4 + 1] | ac[k * 4 + 2] | ac[k * 4 + 3]) mask |= 1u << k;
      if (mask == 0) mask = 1u;                       // keep >= 1 step per tile (an all-zero gather)
      if (tid == 0) {
        mask_s[ti % MASK_RING] = mask;
        mbar_arrive(mask_bar0 + 8 * (ti % MASK_RING));
      }
      for (int k = 0; k < p.koff; ++k) {
        if (!((mask >> k) & 1u)) continue;
        const float* wk = p.w + (size_t)k * NSPLIT * p.n_pad * p.cin_pad;
        for (int c = 0; c < nchunk; ++c, ++g) {
          const int s = g % STAGES;
          const uint32_t ph = (uint32_t)(g / STAGES) & 1u;
          WAIT(empty_bar0 + 8 * s, ph ^ 1u);
          const uint32_t a_dst = smem_u32(a_s + s * a_bytes);
          const uint32_t b_dst = smem_u32(b_s + s * b_bytes);
          // ---- A: 128 rows x 8 chunks of 16 B; 8 lanes cover one row (4 full 128 B lines per warp request).
          // The producers are instruction-issue bound, so the per-copy work is pared down: one 16-byte shared load
          // brings the thread's 4 rulebook entries, a missing row becomes a zero-size copy from row 0 (no pointer
          // selects), row offsets are 32x32->64-bit multiplies, destination offsets are per-thread constants.
          const int col = c * KCH + ch * 4;
          if (col < p.cin_pad && !(p.debug_skip & 64)) {           // columns >= cin_pad are never read by the MMA
            const bool col_ok = (col < cin) && !(p.debug_skip & 1);
            const bool first = col < p.c0;
            const float* abase = first ? (p.in0 + col) : (p.in1 + (col - p.c0));
            const uint32_t ald = first ? (uint32_t)p.ld0 : (uint32_t)p.ld1;
            const int4 j4 = *reinterpret_cast<const int4*>(nb + k * TILE_M + rsub * 4);
            const int jr[4] = {j4.x, j4.y, j4.z, j4.w};
#pragma unroll
            for (int it = 0; it < A_IT; ++it) {
              const uint32_t jc = (uint32_t)max(jr[it], 0);
              cp_async16(a_dst + soff[it], abase + (size_t)jc * ald, (jr[it] >= 0 && col_ok) ? 16u : 0u);
            }
            // ---- B: n_pad rows x 8 chunks (x2 when split): row it*32 + rsub, chunk ch -> same swizzled offsets
            if (!(p.debug_skip & 2)) {
              const float* bsrc = wk + (size_t)rsub * p.cin_pad + col;
#pragma unroll 1
              for (int it = 0; it < nbg; ++it, bsrc += wstride) {
                if ((bmask >> it) & 1u) {
                  cp_async16(b_dst + soff[it], bsrc, 16u);
                  if (SPLIT) cp_async16(b_dst + b_half + soff[it], bsrc + wsplit, 16u);
                }
              }
            }
          }
          // the mbarrier is signalled by the hardware once this thread's copies above have landed
          if (p.debug_skip & 128) mbar_arrive(land_bar0 + 8 * s); else
          cp_async_mbar_arrive_noinc(land_bar0 + 8 * s);
        }
      }
    }
  } else if (warp < MMA_WARP) {
    // =========================== splitters / publishers ===========================
    const int row = tid - SPLIT_WARP0 * 32;           // tile row owned by this thread
    int g = 0, ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      WAIT(mask_bar0 + 8 * (ti % MASK_RING), (uint32_t)(ti / MASK_RING) & 1u);
      const uint32_t mask = *(volatile uint32_t*)&mask_s[ti % MASK_RING];
      const int nst = __popc(mask) * nchunk;
      for (int st = 0; st < nst; ++st, ++g) {
        const int s = g % STAGES;
        WAIT(land_bar0 + 8 * s, (uint32_t)(g / STAGES) & 1u);
        if (SPLIT && !(p.debug_skip & 16)) {
          uint8_t* a_raw = a_s + s * a_bytes;
#pragma unroll
          for (int cch = 0; cch < 8; ++cch) {
            const uint32_t off = sw128(row, cch);
            float4 v = *reinterpret_cast<const float4*>(a_raw + off);
            v.x -= __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
            v.y -= __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
            v.z -= __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
            v.w -= __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
            *reinterpret_cast<float4*>(a_raw + a_half + off) = v;
          }
        }
        if (!(p.debug_skip & 32)) fence_proxy_async_smem();                     // generic-proxy writes (cp.async / st.shared) -> async proxy (MMA)
        mbar_arrive(full_bar0 + 8 * s);
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc_tf32((uint32_t)p.n_pad);
    int g = 0, ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1;
      WAIT(mask_bar0 + 8 * (ti % MASK_RING), (uint32_t)(ti / MASK_RING) & 1u);
      const uint32_t mask = *(volatile uint32_t*)&mask_s[ti % MASK_RING];
      WAIT(acce_bar0 + 8 * buf, ((uint32_t)(ti >> 1) & 1u) ^ 1u);     // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(buf * p.n_pad);
      int nst = 0;
      for (int k = 0; k < p.koff; ++k)
        if ((mask >> k) & 1u) nst += nchunk;
      int st = 0;
      for (int k = 0; k < p.koff; ++k) {
        if (!((mask >> k) & 1u)) continue;
        for (int c = 0; c < nchunk; ++c, ++g, ++st) {
          const int s = g % STAGES;
          const uint32_t ph = (uint32_t)(g / STAGES) & 1u;
          WAIT(full_bar0 + 8 * s, ph);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t adesc = make_desc_k_sw128(smem_u32(a_s + s * a_bytes));
            const uint64_t bdesc = make_desc_k_sw128(smem_u32(b_s + s * b_bytes));
            const uint64_t alo = make_desc_k_sw128(smem_u32(a_s + s * a_bytes + a_half));
            const uint64_t blo = make_desc_k_sw128(smem_u32(b_s + s * b_bytes + b_half));
            const int kc = min(KCH, p.cin_pad - c * KCH);  // multiple of 8
            for (int kk = 0; kk < ((p.debug_skip & 4) ? 0 : kc / 8); ++kk) {
              // advance 8 tf32 = 32 B inside the 128 B swizzle row: +2 in the >>4 address field
              const uint64_t o = (uint64_t)(kk * 2);
              umma_tf32(tacc, adesc + o, bdesc + o, idesc, (st > 0 || kk > 0) ? 1u : 0u);
              if (SPLIT) {
                umma_tf32(tacc, adesc + o, blo + o, idesc, 1u);
                umma_tf32(tacc, alo + o, bdesc + o, idesc, 1u);
              }
            }
            umma_commit(empty_bar0 + 8 * s);
            if (st == nst - 1) umma_commit(accf_bar0 + 8 * buf);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =========================== epilogue ===========================
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    const int et = q * 32 + lane;                      // tile row owned by this thread
    for (int c = et; c < COLV; c += 128) {
      const bool in = c < p.cout;
      colv[c] = (in && p.scale) ? __ldg(p.scale + c) : 1.f;
      colv[COLV + c] = (in && p.shift) ? __ldg(p.shift + c) : 0.f;
      colv[2 * COLV + c] = (in && p.n_ln > 0) ? __ldg(p.ln_g0 + c) : 1.f;
      colv[3 * COLV + c] = (in && p.n_ln > 0) ? __ldg(p.ln_b0 + c) : 0.f;
      colv[4 * COLV + c] = (in && p.n_ln > 1) ? __ldg(p.ln_g1 + c) : 1.f;
      colv[5 * COLV + c] = (in && p.n_ln > 1) ? __ldg(p.ln_b1 + c) : 0.f;
    }
    bar_sync_epilogue();
    int ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1;
      WAIT(accf_bar0 + 8 * buf, (uint32_t)(ti >> 1) & 1u);
      tc_fence_after();
      const uint32_t trow = tmem_base + (uint32_t)(buf * p.n_pad) + ((uint32_t)(q * 32) << 16);
      if (!(p.debug_skip & 8)) epilogue_tile(p, trow, tile * TILE_M, et, colv, stg);
      tc_fence_before();
      mbar_arrive(acce_bar0 + 8 * buf);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, tmem_cols);
}

static size_t smem_bytes_for(int stages, int n_pad, int koff, int nsplit) {
  size_t b = 1024;  // alignment slack
  b += (size_t)stages * nsplit * (a_stage_bytes() + b_stage_bytes(n_pad));
  b += (size_t)2 * koff * TILE_M * 4 + (size_t)2 * koff * 16 + MASK_RING * 4;
  b += 8 + (3 * stages + 4 + MASK_RING) * 8 + 16 + 32;
  b += (size_t)(6 * COLV + TILE_M * STG_LD) * 4;
  return b;
}

}  // namespace ls3d

int ls3d_gather_gemm_bf16x3_launch(const ls3d_gemm_args* a, int num_sms, void* stream);   // gather_gemm_bf16x3.cu
int ls3d_gather_gemm_once_launch(const ls3d_gemm_args* a, int num_sms, void* stream);     // gather_gemm_once.cu
int ls3d_gather_gemm_once_fits(const ls3d_gemm_args* a);

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
  if (a->red0 && (!a->red1 || (a->red_c & 3) || (a->ld_red0 & 3) || (a->ld_red1 & 3) || a->cout != a->red_c ||
                  a->epi != LS3D_EPI_LINEAR))
    return LS3D_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int num_sms = ls3d_num_sms();
  if (a->precise == 2 && a->nbr && a->plan_hdr && ls3d_gather_gemm_once_fits(a)) return ls3d_gather_gemm_once_launch(a, num_sms, stream);
  if (a->precise == 2) return ls3d_gather_gemm_bf16x3_launch(a, num_sms, stream);
  const int ntiles = ls3d_div_up(a->m_out, TILE_M);
  const int grid = ntiles < num_sms ? ntiles : num_sms;          // persistent: one CTA per SM
  const int nsplit = a->precise ? 2 : 1;
  int stages = 4;
  while (stages > 2 && smem_bytes_for(stages, a->n_pad, a->koff, nsplit) > 227 * 1024) --stages;
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
