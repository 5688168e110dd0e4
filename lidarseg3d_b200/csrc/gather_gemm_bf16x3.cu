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
//   * a scout warp lands the next tile's rulebook slice with bulk copies and publishes its active-offset mask.
//
// Persistent CTA, one per SM, 19 warps:
//   warps 0-3   producers : cp.async row gathers (8 lanes = one 128-byte row segment; zero-size copy for missing rows),
//                           one warp per step, four steps in flight
//   warps 4-11  splitters : two groups of four (one thread per tile row = TMEM lane), alternate steps
//   warps 12-15 epilogue  : gemm_epilogue.cuh
//   warp  16    MMA issuer, warp 17 W loader (one elected thread each), warp 18 rulebook scout (one tile ahead)
#include <cuda_bf16.h>

#include "gemm_epilogue.cuh"

namespace ls3d {
namespace bf16x3 {

constexpr int SPLIT_WARP0 = 4;
constexpr int EPI_WARP0 = 12;
constexpr int MMA_WARP = 16;
constexpr int W_WARP = 17;
constexpr int SCOUT_WARP = 18;
constexpr int N_THREADS = 19 * 32;
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

// Development instrumentation (nvcc -DLS3D_PROF): per-CTA cycle counters of where each role waits.
#ifdef LS3D_PROF
__device__ unsigned long long g_prof[148 * 16];
#define PROF_DECL unsigned long long pt0_ = 0, pacc_[6] = {0, 0, 0, 0, 0, 0}
#define PROF_T0 pt0_ = clock64()
#define PROF_ADD(i) pacc_[i] += clock64() - pt0_
#define PROF_DUMP(base, n, cond)                                                         \
  if (cond)                                                                              \
    for (int i_ = 0; i_ < n; ++i_) g_prof[blockIdx.x * 16 + base + i_] = pacc_[i_]
#else
#define PROF_DECL
#define PROF_T0
#define PROF_ADD(i)
#define PROF_DUMP(base, n, cond)
#endif

struct Cfg {
  int rs, ts;           // ring depths: raw smem stages; steps in flight behind them (TMEM A slot + W smem stage each)
  int a_col0;           // first TMEM column of the A slots
  int stack;            // 1: W chunk = [W_hi ; W_lo] stacked along N (2 MMAs per K slice), accumulator = two partial tiles
  int acc_stride;       // TMEM columns per accumulator buffer
};

__global__ void __launch_bounds__(N_THREADS, 1) gather_gemm_bf16x3_kernel(const ls3d_gemm_args p, const Cfg cfg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t w_bytes = (uint32_t)p.n_pad * 128u;  // one [n_pad rows x (32 hi | 32 lo) bf16] W chunk, 128B-swizzled
  uint8_t* raw_s = smem;
  uint8_t* w_s = smem + cfg.rs * RAW_BYTES;
  int* nbr_s = (int*)(w_s + cfg.ts * w_bytes);                      // [2][koff][128]
  uint32_t* mask_s = (uint32_t*)(nbr_s + 2 * p.koff * TILE_M);      // [MASK_RING] active-offset mask per tile
  uint64_t* bars = (uint64_t*)(((uintptr_t)(mask_s + MASK_RING) + 7) & ~(uintptr_t)7);
  uint32_t* tmem_slot = (uint32_t*)(bars + 6 * MAXR + 8 + MASK_RING);
  float* colv = (float*)(((uintptr_t)(tmem_slot + 4) + 15) & ~(uintptr_t)15);
  float* stg = colv + 6 * COLV;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cin = p.c0 + p.c1;
  const int nchunk = (p.cin_pad + KCH - 1) / KCH;
  const int ntiles = (p.m_out + TILE_M - 1) / TILE_M;

  const uint32_t land_bar0 = smem_u32(bars);                    // raw rows landed          [rs]  (32 noinc arrivals)
  const uint32_t rawe_bar0 = smem_u32(bars + MAXR);             // raw stage read           [rs]  (128 splitters)
  const uint32_t afull_bar0 = smem_u32(bars + 2 * MAXR);        // bf16 A slot written      [ts]  (128 splitters)
  const uint32_t aempty_bar0 = smem_u32(bars + 3 * MAXR);       // A slot + W stage consumed [ts] (tcgen05.commit)
  const uint32_t wfull_bar0 = smem_u32(bars + 4 * MAXR);        // W chunk landed           [ts]  (expect_tx)
  const uint32_t accf_bar0 = smem_u32(bars + 6 * MAXR);         // accumulator full  [2]
  const uint32_t acce_bar0 = smem_u32(bars + 6 * MAXR + 2);     // accumulator empty [2]
  const uint32_t nbre_bar0 = smem_u32(bars + 6 * MAXR + 4);     // rulebook buffer free   [2] (4 producer warps)
  const uint32_t nbrl_bar0 = smem_u32(bars + 6 * MAXR + 6);     // rulebook slice landed  [2]
  const uint32_t mask_bar0 = smem_u32(bars + 6 * MAXR + 8);     // rulebook + mask published [MASK_RING]

  constexpr uint32_t TMEM_COLS = 512;
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < MAXR; ++s) {
        mbar_init(land_bar0 + 8 * s, 32);
        mbar_init(rawe_bar0 + 8 * s, 4);             // one arrival per splitter warp: 128 per-thread arrivals on one
        mbar_init(afull_bar0 + 8 * s, 4);            // shared-memory word serialise (~250 cycles per barrier and step)
        mbar_init(aempty_bar0 + 8 * s, 1);
        mbar_init(wfull_bar0 + 8 * s, 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(accf_bar0 + 8 * b, 1);
        mbar_init(acce_bar0 + 8 * b, 4);
        mbar_init(nbre_bar0 + 8 * b, 4);
        mbar_init(nbrl_bar0 + 8 * b, 1);
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
    // Warp w gathers whole steps (global step g with (g & 3) == w): 32 copy instructions (8 lanes = one 128-byte row
    // segment, rows q*32 + it) per barrier wait / arrival, and the four warps work on four different steps at once.
    const int ch = lane & 7;                           // 16-byte chunk of the 128-byte row segment
    const int q = lane >> 3;                           // rows q * 32 + it, it = 0..31
    // swizzled offset of (row q*32 + it, chunk ch) = q*4096 + (it>>3)*1024 + (it&7)*128 + ((ch ^ (it&7)) << 4)
    const uint32_t raw0 = smem_u32(raw_s) + (uint32_t)q * 4096u;
    const char* const base0 = reinterpret_cast<const char*>(p.in0);
    const char* const base1 = reinterpret_cast<const char*>(p.in1);
    const uint32_t ldb0 = (uint32_t)p.ld0 * 4u, ldb1 = (uint32_t)p.ld1 * 4u;
    const int c0 = p.c0, cin_pad = p.cin_pad, koff = p.koff;
    Ring rr(cfg.rs);
    int ti = 0, g = 0;
    PROF_DECL;
#ifdef LS3D_PROF
    const unsigned long long pstart_ = clock64();
#endif
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1;
      const int* nb = nbr_s + buf * koff * TILE_M + q * 32;
      mbar_wait(mask_bar0 + 8 * (ti % MASK_RING), (uint32_t)(ti / MASK_RING) & 1u);   // rulebook tile + mask published
      const uint32_t mask = bcast0(*(volatile uint32_t*)&mask_s[ti % MASK_RING]);
      for (int k = 0; k < koff; ++k) {
        if (!((mask >> k) & 1u)) continue;
        for (int c = 0; c < nchunk; ++c, ++g, rr.next()) {
          if ((g & 3) != warp) continue;
          PROF_T0;
          mbar_wait(rawe_bar0 + 8 * rr.idx, rr.ph ^ 1u);
          PROF_ADD(1);
          PROF_T0;
          const uint32_t a_dst = raw0 + (uint32_t)rr.idx * RAW_BYTES;
          const int col = c * KCH + ch * 4;
          // instruction bound: per 16-byte copy a clamp, a compare and one 32x32+64-bit multiply-add
          if (p.debug_skip & 64) {
          } else if (col < cin && !(p.debug_skip & 1)) {
            const bool first = col < c0;
            const char* src = first ? base0 + (size_t)col * 4 : base1 + (size_t)(col - c0) * 4;
            const uint32_t ldb = first ? ldb0 : ldb1;
#pragma unroll
            for (int i8 = 0; i8 < 4; ++i8) {
              const int4 ja = *reinterpret_cast<const int4*>(nb + k * TILE_M + i8 * 8);
              const int4 jb = *reinterpret_cast<const int4*>(nb + k * TILE_M + i8 * 8 + 4);
              const int jr[8] = {ja.x, ja.y, ja.z, ja.w, jb.x, jb.y, jb.z, jb.w};
#pragma unroll
              for (int i = 0; i < 8; ++i)
                cp_async16(a_dst + i8 * 1024 + i * 128 + ((ch ^ i) << 4),
                           src + (uint64_t)(uint32_t)max(jr[i], 0) * ldb, jr[i] >= 0 ? 16u : 0u);
            }
          } else if (col < cin_pad) {                   // zero columns up to the padded K; beyond it nothing reads
#pragma unroll
            for (int it = 0; it < 32; ++it)
              cp_async16(a_dst + (it >> 3) * 1024 + (it & 7) * 128 + ((ch ^ (it & 7)) << 4), base0, 0u);
          }
          cp_async_mbar_arrive_noinc(land_bar0 + 8 * rr.idx);   // fires once this lane's copies have landed
          PROF_ADD(0);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(nbre_bar0 + 8 * buf);          // this warp no longer reads nbr_s[buf]
    }
#ifdef LS3D_PROF
    pacc_[2] = clock64() - pstart_;
#endif
    PROF_DUMP(0, 3, tid == 0);
  } else if (warp < EPI_WARP0) {
    // =========================== splitters ===========================
    const int grp = (warp - SPLIT_WARP0) >> 2;         // steps with (g & 1) == grp
    const int row = ((warp & 3) << 5) | lane;          // tile row == TMEM lane
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    Ring rr(cfg.rs), tr(cfg.ts);
    int g = 0, ti = 0;
    PROF_DECL;
#ifdef LS3D_PROF
    const unsigned long long pstart_ = clock64();
#endif
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      mbar_wait(mask_bar0 + 8 * (ti % MASK_RING), (uint32_t)(ti / MASK_RING) & 1u);
      const uint32_t mask = *(volatile uint32_t*)&mask_s[ti % MASK_RING];
      const int nst = __popc(mask) * nchunk;
      for (int st = 0; st < nst; ++st, ++g, rr.next(), tr.next()) {
        if ((g & 1) != grp) continue;
        PROF_T0;
        mbar_wait(land_bar0 + 8 * rr.idx, rr.ph);
        PROF_ADD(1);
        PROF_T0;
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
        __syncwarp();
        if (lane == 0) mbar_arrive(rawe_bar0 + 8 * rr.idx);   // the raw stage may be refilled
        PROF_ADD(0);
        PROF_T0;
        mbar_wait(aempty_bar0 + 8 * tr.idx, tr.ph ^ 1u);
        PROF_ADD(2);
        tc_fence_after();
        PROF_T0;
        const uint32_t ta = tmem_a0 + lane_addr + (uint32_t)(tr.idx * A_SLOT_COLS);
        tmem_st16(ta, hi);
        tmem_st16(ta + 16, lo);
        tmem_st_wait();
        PROF_ADD(3);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(afull_bar0 + 8 * tr.idx);
      }
    }
#ifdef LS3D_PROF
    pacc_[4] = clock64() - pstart_;
#endif
    PROF_DUMP(3, 5, row == 0 && grp == 0);
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer ===========================
    // every operand of the issue path is kept provably warp-uniform (kernel parameters, loop counters, bcast0 of loaded
    // values) so that ptxas emits straight-line UTCHMMA / UTCBAR on the uniform datapath instead of per-instruction
    // "elect / R2UR.BROADCAST / branch" waterfall loops (~100 cycles per MMA, measured on the tf32 kernel)
    const uint32_t idesc = make_idesc_bf16((uint32_t)p.n_pad);
    const uint32_t idesc2 = make_idesc_bf16(2u * (uint32_t)p.n_pad);
    const uint32_t tbase = bcast0(tmem_base);
    const uint32_t ta0 = tbase + (uint32_t)cfg.a_col0;
    const uint32_t ws0 = smem_u32(w_s);
    Ring tr(cfg.ts);                                   // A slot (TMEM) and W stage (smem) advance together
    int ti = 0;
    PROF_DECL;
#ifdef LS3D_PROF
    const unsigned long long pstart_ = clock64();
#endif
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1;
      mbar_wait(mask_bar0 + 8 * (ti % MASK_RING), (uint32_t)(ti / MASK_RING) & 1u);
      const uint32_t mask = bcast0(*(volatile uint32_t*)&mask_s[ti % MASK_RING]);
      PROF_T0;
      mbar_wait(acce_bar0 + 8 * buf, ((uint32_t)(ti >> 1) & 1u) ^ 1u);     // epilogue drained this accumulator
      PROF_ADD(1);
      tc_fence_after();
      const uint32_t tacc = tbase + (uint32_t)(buf * cfg.acc_stride);
      const int nst = __popc(mask) * nchunk;
      int st = 0;
      for (int k = 0; k < p.koff; ++k) {
        if (!((mask >> k) & 1u)) continue;
        for (int c = 0; c < nchunk; ++c, ++st, tr.next()) {
          PROF_T0;
          mbar_wait(afull_bar0 + 8 * tr.idx, tr.ph);
          PROF_ADD(2);
          PROF_T0;
          mbar_wait(wfull_bar0 + 8 * tr.idx, tr.ph);
          PROF_ADD(3);
          tc_fence_after();
          PROF_T0;
          const uint64_t bdesc = cfg.stack ? make_desc_k_sw64(ws0 + (uint32_t)tr.idx * w_bytes)
                                           : make_desc_k_sw128(ws0 + (uint32_t)tr.idx * w_bytes);
          const uint32_t a_hi = ta0 + (uint32_t)(tr.idx * A_SLOT_COLS);
          const uint32_t a_lo = a_hi + 16;
          const int nsl = ((p.debug_skip & 4) ? 0 : min(KCH, p.cin_pad - c * KCH) / 16);
          if (elect_one()) {
            if (cfg.stack) {
              // W chunk rows = [W_hi (n_pad rows) ; W_lo (n_pad rows)] x 32 K (64-byte rows): x_hi meets both in one
              // MMA of N = 2 n_pad (columns [0,n) += x_hi.W_hi, [n,2n) += x_hi.W_lo), x_lo meets W_hi in a second one.
              // The A operand costs ~64+ cycles per instruction whatever N is, so fewer, wider MMAs win.
              for (int j = 0; j < nsl; ++j) {
                const uint64_t o = (uint64_t)(2 * j);       // 16 bf16 = 32 bytes = +2 in the (address >> 4) field
                umma_bf16_ts(tacc, a_hi + 8 * j, bdesc + o, idesc2, (st > 0 || j > 0) ? 1u : 0u);
                umma_bf16_ts(tacc, a_lo + 8 * j, bdesc + o, idesc, 1u);
              }
            } else {
              for (int j = 0; j < nsl; ++j) {
                // W row = [hi 0..31 | lo 0..31] (128-byte rows)
                const uint64_t o = (uint64_t)(2 * j);
                umma_bf16_ts(tacc, a_hi + 8 * j, bdesc + o, idesc, (st > 0 || j > 0) ? 1u : 0u);
                umma_bf16_ts(tacc, a_hi + 8 * j, bdesc + 4 + o, idesc, 1u);
                umma_bf16_ts(tacc, a_lo + 8 * j, bdesc + o, idesc, 1u);
              }
            }
            umma_commit(aempty_bar0 + 8 * tr.idx);           // frees the A slot and the W stage of this step
            if (st == nst - 1) umma_commit(accf_bar0 + 8 * buf);
          }
          __syncwarp();
          PROF_ADD(0);
        }
      }
    }
#ifdef LS3D_PROF
    pacc_[4] = clock64() - pstart_;
#endif
    PROF_DUMP(8, 5, lane == 0);
  } else if (warp == W_WARP) {
    // =========================== W loader ===========================
    if (lane == 0) {
      Ring wr(cfg.ts);
      int ti = 0;
      const uint8_t* wg = reinterpret_cast<const uint8_t*>(p.w);
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
        mbar_wait(mask_bar0 + 8 * (ti % MASK_RING), (uint32_t)(ti / MASK_RING) & 1u);
        const uint32_t mask = *(volatile uint32_t*)&mask_s[ti % MASK_RING];
        for (int k = 0; k < p.koff; ++k) {
          if (!((mask >> k) & 1u)) continue;
          for (int c = 0; c < nchunk; ++c, wr.next()) {
            mbar_wait(aempty_bar0 + 8 * wr.idx, wr.ph ^ 1u);
            mbar_arrive_expect_tx(wfull_bar0 + 8 * wr.idx, w_bytes);
            bulk_g2s(smem_u32(w_s + wr.idx * w_bytes), wg + ((size_t)k * nchunk + c) * w_bytes, w_bytes,
                     wfull_bar0 + 8 * wr.idx);
          }
        }
      }
    }
  } else if (warp == SCOUT_WARP) {
    // =========================== rulebook scout ===========================
    // Runs one tile ahead: lands the tile's [koff][128] rulebook slice in shared memory (one 512-byte bulk copy per offset
    // when the layout allows it), derives the mask of offsets that have at least one neighbour in the tile (the others
    // are skipped by every role) and publishes both with one mbarrier arrival.
    const bool bulk_ok = p.nbr && (p.m_out & 3) == 0 && ((uintptr_t)p.nbr & 15) == 0;
    int ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1;
      int* nb = nbr_s + buf * p.koff * TILE_M;
      mbar_wait(nbre_bar0 + 8 * buf, ((uint32_t)(ti >> 1) & 1u) ^ 1u);       // producers done with this buffer
      const int r0 = tile * TILE_M;
      if (bulk_ok && r0 + TILE_M <= p.m_out) {
        if (lane == 0) {
          mbar_arrive_expect_tx(nbrl_bar0 + 8 * buf, (uint32_t)p.koff * 512u);
          for (int k = 0; k < p.koff; ++k)
            bulk_g2s(smem_u32(nb + k * TILE_M), p.nbr + (size_t)k * p.m_out + r0, 512u, nbrl_bar0 + 8 * buf);
        }
      } else {
        for (int k = 0; k < p.koff; ++k)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = r0 + i * 32 + lane;
            nb[k * TILE_M + i * 32 + lane] = (r < p.m_out) ? (p.nbr ? __ldg(p.nbr + (size_t)k * p.m_out + r) : r) : -1;
          }
        __syncwarp();
        if (lane == 0) mbar_arrive(nbrl_bar0 + 8 * buf);
      }
      mbar_wait(nbrl_bar0 + 8 * buf, (uint32_t)(ti >> 1) & 1u);
      uint32_t mask = 0;
      for (int k = 0; k < p.koff; ++k) {
        const int4 v = *reinterpret_cast<const int4*>(nb + k * TILE_M + lane * 4);
        if (__ballot_sync(0xffffffffu, (v.x & v.y & v.z & v.w) >= 0)) mask |= 1u << k;
      }
      if (mask == 0) mask = 1u;                          // keep >= 1 step per tile (an all-zero gather)
      if (lane == 0) {
        mask_s[ti % MASK_RING] = mask;
        mbar_arrive(mask_bar0 + 8 * (ti % MASK_RING));
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
    PROF_DECL;
#ifdef LS3D_PROF
    const unsigned long long pstart_ = clock64();
#endif
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1;
      PROF_T0;
      mbar_wait(accf_bar0 + 8 * buf, (uint32_t)(ti >> 1) & 1u);
      PROF_ADD(0);
      tc_fence_after();
      const uint32_t trow = tmem_base + (uint32_t)(buf * cfg.acc_stride) + ((uint32_t)(q * 32) << 16);
      if (!(p.debug_skip & 8)) epilogue_tile(p, trow, tile * TILE_M, et, colv, stg, cfg.stack ? (uint32_t)p.n_pad : 0u);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce_bar0 + 8 * buf);
    }
#ifdef LS3D_PROF
    pacc_[1] = clock64() - pstart_;
#endif
    PROF_DUMP(13, 2, et == 0);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

static size_t smem_bytes_for(int rs, int ws, int n_pad, int koff) {
  size_t b = 1024;  // alignment slack
  b += (size_t)rs * RAW_BYTES + (size_t)ws * n_pad * 128;
  b += (size_t)2 * koff * TILE_M * 4 + MASK_RING * 4;
  b += 8 + (6 * MAXR + 8 + MASK_RING) * 8 + 16 + 32;
  b += (size_t)(6 * COLV + TILE_M * STG_LD) * 4;
  return b;
}

}  // namespace bf16x3

// ------------------------------------------------------------------------------------------------------------------
// Weight image of the bf16x3 engines (ls3d_gemm_args.w with precise = 2; also the chunks ls3d_sffm_decoder streams).
// One block of n_pad * 128 bytes per (offset, 32-channel chunk), stored as the swizzled shared-memory image so that a single
// cp.async.bulk lands a ready tcgen05 K-major B tile:
//   n_pad <= 96 ("stacked"): 2 n_pad rows of 64 bytes; rows [0, n) = bf16 hi of W[n, 32c .. 32c+31], rows [n, 2n) = bf16 lo;
//                SWIZZLE_64B: the 16-byte unit u of row r holds logical unit u ^ ((r >> 1) & 3);
//   n_pad  > 96 : n_pad rows of 128 bytes = [hi (32) | lo (32)]; SWIZZLE_128B: unit u of row r holds logical unit u ^ (r & 7).
// hi = bf16_rn(w), lo = bf16_rn(w - hi).
__global__ void pack_bf16x3_kernel(const float* __restrict__ w_kio, int koff, int cin, int cout, int n_pad, int nchunk,
                                   __nv_bfloat16* __restrict__ out) {
  const bool stacked = n_pad <= 96;
  const int rows = stacked ? 2 * n_pad : n_pad, units = stacked ? 4 : 8;
  const long long total = (long long)koff * nchunk * rows * units * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i & 7);
    long long t = i >> 3;
    const int u = (int)(t % units);
    t /= units;
    const int r = (int)(t % rows);
    t /= rows;
    const int c = (int)(t % nchunk), k = (int)(t / nchunk);
    const int lu = stacked ? (u ^ ((r >> 1) & 3)) : (u ^ (r & 7));
    const bool lo = stacked ? (r >= n_pad) : (lu >= 4);
    const int n = stacked ? (r >= n_pad ? r - n_pad : r) : r;
    const int ch = 32 * c + 8 * (lu & 3) + e;
    float v = 0.f;
    if (n < cout && ch < cin) v = w_kio[((size_t)k * cin + ch) * cout + n];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[i] = lo ? __float2bfloat16_rn(v - __bfloat162float(h)) : h;
  }
}

}  // namespace ls3d (reopened below)

extern "C" int ls3d_gemm_packed_bytes(int32_t koff, int32_t cin, int32_t cout, int64_t* bytes, int32_t* cin_pad, int32_t* n_pad) {
  if (!bytes || koff < 1 || cin < 1 || cout < 1 || cout > 256) return LS3D_ERR_ARG;
  const int cp = (cin + 15) / 16 * 16, np = (cout + 15) / 16 * 16;
  *bytes = (int64_t)koff * ((cp + 31) / 32) * np * 128;
  if (cin_pad) *cin_pad = cp;
  if (n_pad) *n_pad = np;
  return LS3D_OK;
}

// w_kio: fp32 [koff][cin][cout] (spconv weight [kz, ky, kx, Cin, Cout] flattened over the offsets; an nn.Linear weight is its
// transpose with koff = 1) on the device; out: ls3d_gemm_packed_bytes bytes, 16-byte aligned
extern "C" int ls3d_gemm_pack_bf16x3(const float* w_kio, int32_t koff, int32_t cin, int32_t cout, void* out, void* stream) {
  int64_t bytes;
  int32_t cp, np;
  if (!w_kio || !out || (((uintptr_t)out) & 15) || ls3d_gemm_packed_bytes(koff, cin, cout, &bytes, &cp, &np)) return LS3D_ERR_ARG;
  const long long total = bytes / 2;
  const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  ls3d::pack_bf16x3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w_kio, koff, cin, cout, np, (cp + 31) / 32,
                                                                  (__nv_bfloat16*)out);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

namespace ls3d {

}  // namespace ls3d

#ifdef LS3D_PROF
extern "C" int ls3d_debug_gemm_prof(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, ls3d::bf16x3::g_prof, sizeof(unsigned long long) * 148 * 16);
}
#endif

// called by ls3d_gather_gemm (gather_gemm.cu) for args->precise == 2, after the common argument checks
int ls3d_gather_gemm_bf16x3_launch(const ls3d_gemm_args* a, int num_sms, void* stream) {
  using namespace ls3d;
  using namespace ls3d::bf16x3;
  if (a->cin_pad % 16 || a->n_pad > 192) return LS3D_ERR_ARG;
  Cfg cfg;
  cfg.stack = a->n_pad <= 96 ? 1 : 0;          // must match PackedWeight._pack_bf16x3 (lidarseg3d_b200/gemm.py)
  cfg.acc_stride = cfg.stack ? 2 * a->n_pad : a->n_pad;
  const int ts_max = 2 * cfg.acc_stride <= 256 ? 8 : 4;         // TMEM columns left for the A slots
  cfg.a_col0 = 2 * cfg.acc_stride <= 256 ? 256 : 384;
  // deepest (A slot + W stage) ring that leaves room for >= 4 raw stages, then as many raw stages as fit (<= 8)
  cfg.ts = ts_max;
  while (cfg.ts > 2 && smem_bytes_for(4, cfg.ts, a->n_pad, a->koff) > 227 * 1024) --cfg.ts;
  cfg.rs = 8;
  while (cfg.rs > 2 && smem_bytes_for(cfg.rs, cfg.ts, a->n_pad, a->koff) > 227 * 1024) --cfg.rs;
  const size_t smem = smem_bytes_for(cfg.rs, cfg.ts, a->n_pad, a->koff);
  if (smem > 227 * 1024) return LS3D_ERR_ARG;
  const int ntiles = ls3d_div_up(a->m_out, TILE_M);
  const int grid = ntiles < num_sms ? ntiles : num_sms;          // persistent: one CTA per SM
  static bool optin[64] = {false};
  cudaError_t e = ls3d_optin_smem(gather_gemm_bf16x3_kernel, optin);
  if (e != cudaSuccess) return (int)e;
  gather_gemm_bf16x3_kernel<<<grid, N_THREADS, smem, (cudaStream_t)stream>>>(*a, cfg);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}
