// Gather-GEMM, "bf16x3" engine (precise = 2): fp32 activations and weights are split on the fly into bf16 hi + lo parts
// and every product is formed as  x_hi.W_hi + x_hi.W_lo + x_lo.W_hi  on the tcgen05 tensor cores (kind::f16, fp32
// accumulate in TMEM): ~2^-17 relative error per product (40x tighter than single-pass TF32) at half the tensor-pipe time
// of 3xTF32.  Same contract / epilogues as gather_gemm.cu (see there for the reference call sites it replaces:
// det3d/models/backbones/scn_unet.py:15-20,39-46,89-160 and the Linear layers of det3d/models/point_heads/*.py).
//
// What is different from the tf32 kernel (measured there: the smem-descriptor A operand costs ~64 cycles per MMA whatever
// N is, and a 2-4 stage ring with a splitter hop in it is latency bound):
//   * the A operand lives in TENSOR MEMORY: splitter warps read the landed fp32 rows from shared memory, split them in
//     registers and write bf16 hi / lo with tcgen05.st; tcgen05.mma reads A from TMEM and only W from shared memory;
//   * three decoupled rings: raw gathered rows (smem, freed as soon as they are split), bf16 A slots (TMEM, 4-8 deep,
//     freed by tcgen05.commit), W chunks (smem, one cp.async.bulk per step from a pre-swizzled global image);
//   * the next tile's rulebook rows are prefetched into registers while the current tile's gathers are issued.
//
// Persistent CTA, one per SM, 18 warps:
//   warps 0-3   producers : cp.async row gathers (8 lanes = one 128-byte row segment; zero-size copy for missing rows)
//   warps 4-11  splitters : two groups of four (one thread per tile row = TMEM lane), alternate steps
//   warps 12-15 epilogue  : gemm_epilogue.cuh
//   warp  16    MMA issuer, warp 17 W loader (one elected thread each)
#include "gemm_epilogue.cuh"

namespace ls3d {
namespace bf16x3 {

constexpr int N_PROD = 128;
constexpr int SPLIT_WARP0 = 4;
constexpr int EPI_WARP0 = 12;
constexpr int MMA_WARP = 16;
constexpr int W_WARP = 17;
constexpr int N_THREADS = 18 * 32;
constexpr int RAW_BYTES = TILE_M * 128;       // one [128 rows x 32 fp32] gathered chunk
constexpr int MAXR = 8;                       // max depth of each ring
constexpr int A_SLOT_COLS = 32;               // TMEM columns per A slot: 16 (hi, 32 bf16) + 16 (lo)

struct Ring {
  int n;
  int idx;
  uint32_t ph;
  __device__ __forceinline__ Ring(int n_) : n(n_), idx(0), ph(0) {}
  __device__ __forceinline__ void next() {
    if (++idx == n) {
      idx = 0;
      ph ^= 1u;
    }
  }
};

__device__ __forceinline__ void bar_sync_producers() { asm volatile("bar.sync 1, %0;" ::"n"(N_PROD) : "memory"); }

struct Cfg {
  int rs, ws, ts;       // ring depths: raw smem stages, W smem stages, TMEM A slots
  int a_col0;           // first TMEM column of the A slots
};

__global__ void __launch_bounds__(N_THREADS, 1) gather_gemm_bf16x3_kernel(const ls3d_gemm_args p, const Cfg cfg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t w_bytes = (uint32_t)p.n_pad * 128u;  // one [n_pad rows x (32 hi | 32 lo) bf16] W chunk, 128B-swizzled
  uint8_t* raw_s = smem;
  uint8_t* w_s = smem + cfg.rs * RAW_BYTES;
  int* nbr_s = (int*)(w_s + cfg.ws * w_bytes);                      // [2][koff][128]
  uint32_t* wmask_s = (uint32_t*)(nbr_s + 2 * p.koff * TILE_M);     // [2][4] per-warp active-offset masks
  uint32_t* mask_s = wmask_s + 8;                                   // [MASK_RING]
  uint64_t* bars = (uint64_t*)(((uintptr_t)(mask_s + MASK_RING) + 7) & ~(uintptr_t)7);
  uint32_t* tmem_slot = (uint32_t*)(bars + 6 * MAXR + 4 + MASK_RING);
  float* colv = (float*)(((uintptr_t)(tmem_slot + 4) + 15) & ~(uintptr_t)15);
  float* stg = colv + 6 * COLV;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cin = p.c0 + p.c1;
  const int nchunk = (p.cin_pad + KCH - 1) / KCH;
  const int ntiles = (p.m_out + TILE_M - 1) / TILE_M;

  const uint32_t land_bar0 = smem_u32(bars);                    // raw rows landed          [rs]  (128 noinc arrivals)
  const uint32_t rawe_bar0 = smem_u32(bars + MAXR);             // raw stage read           [rs]  (128 splitters)
  const uint32_t afull_bar0 = smem_u32(bars + 2 * MAXR);        // bf16 A slot written      [ts]  (128 splitters)
  const uint32_t aempty_bar0 = smem_u32(bars + 3 * MAXR);       // A slot consumed          [ts]  (tcgen05.commit)
  const uint32_t wfull_bar0 = smem_u32(bars + 4 * MAXR);        // W chunk landed           [ws]  (expect_tx)
  const uint32_t wempty_bar0 = smem_u32(bars + 5 * MAXR);       // W chunk consumed         [ws]  (tcgen05.commit)
  const uint32_t accf_bar0 = smem_u32(bars + 6 * MAXR);         // accumulator full  [2]
  const uint32_t acce_bar0 = smem_u32(bars + 6 * MAXR + 2);     // accumulator empty [2]
  const uint32_t mask_bar0 = smem_u32(bars + 6 * MAXR + 4);     // mask published    [MASK_RING]

  constexpr uint32_t TMEM_COLS = 512;
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < MAXR; ++s) {
        mbar_init(land_bar0 + 8 * s, N_PROD);
        mbar_init(rawe_bar0 + 8 * s, 128);
        mbar_init(afull_bar0 + 8 * s, 128);
        mbar_init(aempty_bar0 + 8 * s, 1);
        mbar_init(wfull_bar0 + 8 * s, 1);
        mbar_init(wempty_bar0 + 8 * s, 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(accf_bar0 + 8 * b, 1);
        mbar_init(acce_bar0 + 8 * b, 128);
      }
      for (int m = 0; m < MASK_RING; ++m) mbar_init(mask_bar0 + 8 * m, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a0 = tmem_base + (uint32_t)cfg.a_col0;

  if (warp < SPLIT_WARP0) {
    // =========================== producers ===========================
    const int ch = tid & 7;                            // 16-byte chunk of the 128-byte row segment
    const int rsub = tid >> 3;                         // rows rsub + 16 * it, it = 0..7
    uint32_t soff[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) soff[it] = sw128(rsub + 16 * it, ch);
    int jv[MAX_KOFF];                                  // rulebook column of tile row `tid`, all offsets
    auto load_rulebook = [&](int tile) {
      const int r = tile * TILE_M + tid;
#pragma unroll
      for (int k = 0; k < MAX_KOFF; ++k) {
        jv[k] = -1;
        if (k < p.koff && r < p.m_out) jv[k] = p.nbr ? __ldg(p.nbr + (size_t)k * p.m_out + r) : r;
      }
    };
    if ((int)blockIdx.x < ntiles) load_rulebook(blockIdx.x);
    Ring rr(cfg.rs);
    int ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1;
      int* nb = nbr_s + buf * p.koff * TILE_M;
      uint32_t wm = 0;
#pragma unroll
      for (int k = 0; k < MAX_KOFF; ++k) {
        if (k < p.koff) {
          nb[k * TILE_M + (tid & 15) * 8 + (tid >> 4)] = jv[k];
          if (__ballot_sync(0xffffffffu, jv[k] >= 0)) wm |= 1u << k;
        }
      }
      if (lane == 0) wmask_s[buf * 4 + warp] = wm;
      bar_sync_producers();      // nbr_s[buf] complete; also: everyone finished the tile (ti-1) that used nbr_s[buf^1]
      uint32_t mask = wmask_s[buf * 4] | wmask_s[buf * 4 + 1] | wmask_s[buf * 4 + 2] | wmask_s[buf * 4 + 3];
      if (mask == 0) mask = 1u;                        // keep >= 1 step per tile (an all-zero gather)
      if (tid == 0) {
        mask_s[ti % MASK_RING] = mask;
        mbar_arrive(mask_bar0 + 8 * (ti % MASK_RING));
      }
      if (tile + (int)gridDim.x < ntiles) load_rulebook(tile + gridDim.x);   // in flight while this tile is issued
      for (int k = 0; k < p.koff; ++k) {
        if (!((mask >> k) & 1u)) continue;
        const int4 ja = *reinterpret_cast<const int4*>(nb + k * TILE_M + rsub * 8);
        const int4 jb = *reinterpret_cast<const int4*>(nb + k * TILE_M + rsub * 8 + 4);
        const int jr[8] = {ja.x, ja.y, ja.z, ja.w, jb.x, jb.y, jb.z, jb.w};
        for (int c = 0; c < nchunk; ++c) {
          mbar_wait(rawe_bar0 + 8 * rr.idx, rr.ph ^ 1u);
          const uint32_t a_dst = smem_u32(raw_s + rr.idx * RAW_BYTES);
          const int col = c * KCH + ch * 4;
          if (col < p.cin_pad) {                        // columns >= cin_pad are never read by the MMA
            const bool col_ok = (col < cin) && !(p.debug_skip & 1);
            const bool first = col < p.c0;
            const float* abase = first ? (p.in0 + col) : (p.in1 + (col - p.c0));
            const uint32_t ald = first ? (uint32_t)p.ld0 : (uint32_t)p.ld1;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const uint32_t jc = (uint32_t)max(jr[it], 0);
              cp_async16(a_dst + soff[it], abase + (size_t)jc * ald, (jr[it] >= 0 && col_ok) ? 16u : 0u);
            }
          }
          cp_async_mbar_arrive_noinc(land_bar0 + 8 * rr.idx);   // fires once this thread's copies have landed
          rr.next();
        }
      }
    }
  } else if (warp < EPI_WARP0) {
    // =========================== splitters ===========================
    const int grp = (warp - SPLIT_WARP0) >> 2;         // steps with (g & 1) == grp
    const int row = ((warp & 3) << 5) | lane;          // tile row == TMEM lane
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    Ring rr(cfg.rs), tr(cfg.ts);
    int g = 0, ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      mbar_wait(mask_bar0 + 8 * (ti % MASK_RING), (uint32_t)(ti / MASK_RING) & 1u);
      const uint32_t mask = *(volatile uint32_t*)&mask_s[ti % MASK_RING];
      const int nst = __popc(mask) * nchunk;
      for (int st = 0; st < nst; ++st, ++g, rr.next(), tr.next()) {
        if ((g & 1) != grp) continue;
        mbar_wait(land_bar0 + 8 * rr.idx, rr.ph);
        const uint8_t* a_raw = raw_s + rr.idx * RAW_BYTES;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int cch = 0; cch < 8; ++cch) {
          const float4 v = *reinterpret_cast<const float4*>(a_raw + sw128(row, cch));
          const uint32_t h0 = pack_bf16x2(v.x, v.y), h1 = pack_bf16x2(v.z, v.w);
          hi[2 * cch] = h0;
          hi[2 * cch + 1] = h1;
          lo[2 * cch] = pack_bf16x2(v.x - __uint_as_float(h0 << 16), v.y - __uint_as_float(h0 & 0xFFFF0000u));
          lo[2 * cch + 1] = pack_bf16x2(v.z - __uint_as_float(h1 << 16), v.w - __uint_as_float(h1 & 0xFFFF0000u));
        }
        mbar_arrive(rawe_bar0 + 8 * rr.idx);            // the raw stage may be refilled
        mbar_wait(aempty_bar0 + 8 * tr.idx, tr.ph ^ 1u);
        tc_fence_after();
        const uint32_t ta = tmem_a0 + lane_addr + (uint32_t)(tr.idx * A_SLOT_COLS);
        tmem_st16(ta, hi);
        tmem_st16(ta + 16, lo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(afull_bar0 + 8 * tr.idx);
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer ===========================
    // every operand of the issue path is kept provably warp-uniform (kernel parameters, loop counters, bcast0 of loaded
    // values) so that ptxas emits straight-line UTCHMMA / UTCBAR on the uniform datapath instead of per-instruction
    // "elect / R2UR.BROADCAST / branch" waterfall loops (~100 cycles per MMA, measured on the tf32 kernel)
    const uint32_t idesc = make_idesc_bf16((uint32_t)p.n_pad);
    const uint32_t tbase = bcast0(tmem_base);
    const uint32_t ta0 = tbase + (uint32_t)cfg.a_col0;
    const uint32_t ws0 = smem_u32(w_s);
    Ring tr(cfg.ts), wr(cfg.ws);
    int ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1;
      mbar_wait(mask_bar0 + 8 * (ti % MASK_RING), (uint32_t)(ti / MASK_RING) & 1u);
      const uint32_t mask = bcast0(*(volatile uint32_t*)&mask_s[ti % MASK_RING]);
      mbar_wait(acce_bar0 + 8 * buf, ((uint32_t)(ti >> 1) & 1u) ^ 1u);     // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tbase + (uint32_t)(buf * p.n_pad);
      const int nst = __popc(mask) * nchunk;
      int st = 0;
      for (int k = 0; k < p.koff; ++k) {
        if (!((mask >> k) & 1u)) continue;
        for (int c = 0; c < nchunk; ++c, ++st, tr.next(), wr.next()) {
          mbar_wait(afull_bar0 + 8 * tr.idx, tr.ph);
          mbar_wait(wfull_bar0 + 8 * wr.idx, wr.ph);
          tc_fence_after();
          const uint64_t bdesc = make_desc_k_sw128(ws0 + (uint32_t)wr.idx * w_bytes);
          const uint32_t a_hi = ta0 + (uint32_t)(tr.idx * A_SLOT_COLS);
          const uint32_t a_lo = a_hi + 16;
          const int nsl = ((p.debug_skip & 4) ? 0 : min(KCH, p.cin_pad - c * KCH) / 16);
          if (elect_one()) {
            for (int j = 0; j < nsl; ++j) {
              // 16 bf16 = 32 bytes = +2 in the descriptor's (address >> 4) field; W row = [hi 0..31 | lo 0..31]
              const uint64_t o = (uint64_t)(2 * j);
              umma_bf16_ts(tacc, a_hi + 8 * j, bdesc + o, idesc, (st > 0 || j > 0) ? 1u : 0u);
              umma_bf16_ts(tacc, a_hi + 8 * j, bdesc + 4 + o, idesc, 1u);
              umma_bf16_ts(tacc, a_lo + 8 * j, bdesc + o, idesc, 1u);
            }
            umma_commit(aempty_bar0 + 8 * tr.idx);
            umma_commit(wempty_bar0 + 8 * wr.idx);
            if (st == nst - 1) umma_commit(accf_bar0 + 8 * buf);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == W_WARP) {
    // =========================== W loader ===========================
    if (lane == 0) {
      Ring wr(cfg.ws);
      int ti = 0;
      const uint8_t* wg = reinterpret_cast<const uint8_t*>(p.w);
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
        mbar_wait(mask_bar0 + 8 * (ti % MASK_RING), (uint32_t)(ti / MASK_RING) & 1u);
        const uint32_t mask = *(volatile uint32_t*)&mask_s[ti % MASK_RING];
        for (int k = 0; k < p.koff; ++k) {
          if (!((mask >> k) & 1u)) continue;
          for (int c = 0; c < nchunk; ++c, wr.next()) {
            mbar_wait(wempty_bar0 + 8 * wr.idx, wr.ph ^ 1u);
            mbar_arrive_expect_tx(wfull_bar0 + 8 * wr.idx, w_bytes);
            bulk_g2s(smem_u32(w_s + wr.idx * w_bytes), wg + ((size_t)k * nchunk + c) * w_bytes, w_bytes,
                     wfull_bar0 + 8 * wr.idx);
          }
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
      mbar_wait(accf_bar0 + 8 * buf, (uint32_t)(ti >> 1) & 1u);
      tc_fence_after();
      const uint32_t trow = tmem_base + (uint32_t)(buf * p.n_pad) + ((uint32_t)(q * 32) << 16);
      if (!(p.debug_skip & 8)) epilogue_tile(p, trow, tile * TILE_M, et, colv, stg);
      tc_fence_before();
      mbar_arrive(acce_bar0 + 8 * buf);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

static size_t smem_bytes_for(int rs, int ws, int n_pad, int koff) {
  size_t b = 1024;  // alignment slack
  b += (size_t)rs * RAW_BYTES + (size_t)ws * n_pad * 128;
  b += (size_t)2 * koff * TILE_M * 4 + 8 * 4 + MASK_RING * 4;
  b += 8 + (6 * MAXR + 4 + MASK_RING) * 8 + 16 + 32;
  b += (size_t)(6 * COLV + TILE_M * STG_LD) * 4;
  return b;
}

}  // namespace bf16x3
}  // namespace ls3d

// called by ls3d_gather_gemm (gather_gemm.cu) for args->precise == 2, after the common argument checks
int ls3d_gather_gemm_bf16x3_launch(const ls3d_gemm_args* a, int num_sms, void* stream) {
  using namespace ls3d;
  using namespace ls3d::bf16x3;
  if (a->cin_pad % 16 || a->n_pad > 192) return LS3D_ERR_ARG;
  Cfg cfg;
  cfg.ts = a->n_pad <= 128 ? 8 : 4;
  cfg.a_col0 = a->n_pad <= 128 ? 256 : 384;
  int d = 6;
  while (d > 2 && smem_bytes_for(d, d, a->n_pad, a->koff) > 227 * 1024) --d;
  cfg.rs = cfg.ws = d;
  const size_t smem = smem_bytes_for(d, d, a->n_pad, a->koff);
  if (smem > 227 * 1024) return LS3D_ERR_ARG;
  const int ntiles = ls3d_div_up(a->m_out, TILE_M);
  const int grid = ntiles < num_sms ? ntiles : num_sms;          // persistent: one CTA per SM
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(gather_gemm_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    smem_set = 227 * 1024;
  }
  gather_gemm_bf16x3_kernel<<<grid, N_THREADS, smem, (cudaStream_t)stream>>>(*a, cfg);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}
