// Sparse convolution, "gather-once" engine: the default kernel behind ls3d_gather_gemm for launches that carry a tile plan.
//
// Same contract, arithmetic (error-compensated bf16x3 products, fp32 accumulate in TMEM) and fused epilogues as
// gather_gemm_bf16x3.cu; what changes is how the A operand reaches the tensor cores.  The per-pair engine fetches one
// 128-byte row segment from L2 per (rulebook pair, 32-channel chunk): 5-15 fetches per output row in LiDAR scenes, and
// that L2->SM gather - not HBM, not the tensor pipe - bounded it (6.2x more on-chip gather traffic than there
// are distinct rows; "No Eligible" 70 %).  Here a rulebook is compiled ONCE per cached table (spconv's indice_key, reused
// by 6-9 convolutions) into a TILE PLAN (ls3d_tile_plan_build, below):
//   per 128-row output tile: the sorted list of DISTINCT input rows any of its kernel offsets touches (<= 512 rows; a tile
//   that touches more is cut into passes over disjoint offset ranges) and a uint16 table local[k][row] = position of the
//   (k, row) neighbour in that list;
// and the convolution kernel
//   * STAGES those distinct rows once per 32-channel chunk: coalesced 128-bit loads -> bf16 hi / lo split in registers
//     (once per distinct row instead of once per pair) -> swizzled shared memory, double buffered;
//   * serves all 27 offsets from that stage: per (offset, chunk) step each tile row copies its neighbour's 128 bytes
//     shared memory -> registers -> tensor memory (tcgen05.st), where the MMA reads the A operand (TS mode);
//   * K loop order = chunk-outer / offset-inner so that a stage is used by every offset before it is replaced.
// Per step the SM now moves 16 KB shared -> tensor memory (~128-220 cycles) against an MMA floor of 96 / 192 / 384 cycles
// for Cout = 32 / 64 / 128 (bf16x3 = 3 products): the convolution is tensor-pipe bound for Cout >= 64.
//
// Reference call sites replaced: spconv indice_conv as used by det3d/models/backbones/scn_unet.py:15-20,39-46,89-160.
//
// Persistent CTA, one per SM, 19 warps:
//   warps 0-3   stagers   : distinct rows -> split -> stage buffer (one (pass, chunk) ahead)
//   warps 4-11  splitters : two groups of four (thread = tile row = TMEM lane), alternate steps: stage -> TMEM A slot
//   warps 12-15 epilogue  : gemm_epilogue.cuh
//   warp  16    MMA issuer, warp 17 W loader (one elected thread each), warp 18 plan loader (one tile ahead)
#include <cstdlib>

#include "gemm_epilogue.cuh"
#include "scan.cuh"

namespace ls3d {
namespace once {

constexpr int S_CAP = 512;                    // distinct rows a pass may stage
constexpr int MAXP = 8;                       // passes per tile (a pass holds >= 4 offsets' worth of rows: 27 / 4 -> 7)
constexpr int HDR_INTS = 32;                  // per tile: n_pass, then MAXP x {kmask, pool base, count}; 128 bytes
constexpr uint32_t NONE16 = 0xFFFFu;
constexpr int STAGE_BYTES = S_CAP * 128;      // [row][hi 32 bf16 | lo 32 bf16], 16-byte units swizzled by (row & 7)

constexpr int SPLIT_WARP0 = 4;
constexpr int EPI_WARP0 = 12;
constexpr int MMA_WARP = 16;
constexpr int W_WARP = 17;
constexpr int PLAN_WARP = 18;
constexpr int N_THREADS = 19 * 32;
constexpr int MAXR = 8;
constexpr int A_SLOT_COLS = 32;

// Development instrumentation (nvcc -DLS3D_PROF): event timeline of CTA 0 - (role, event, counter, clock64) records.
#ifdef LS3D_PROF
constexpr int TRACE_CAP = 2048;
__device__ unsigned long long g_trace[8 * TRACE_CAP];
__device__ unsigned int g_trace_n[8];
#define TRACE(role, ev, idx)                                                                                    \
  do {                                                                                                          \
    if (blockIdx.x == 0 && trc_ < TRACE_CAP) {                                                                  \
      g_trace[(role) * TRACE_CAP + trc_] = ((unsigned long long)(ev) << 56) | ((unsigned long long)((idx) & 0xFFFF) << 40) | \
                                           ((unsigned long long)clock64() & 0xFFFFFFFFFFull);                   \
      g_trace_n[role] = ++trc_;                                                                                 \
    }                                                                                                           \
  } while (0)
#else
#define TRACE(role, ev, idx)
#endif

struct Ring {
  int n, idx;
  uint32_t ph;
  __device__ __forceinline__ Ring(int n_) : n(n_), idx(0), ph(0) {}
  __device__ __forceinline__ void next() {
    if (++idx == n) {
      idx = 0;
      ph ^= 1u;
    }
  }
};

struct Cfg {
  int ts;               // steps in flight (TMEM A slot + W smem stage each)
  int grp;              // offsets per step (A slot = grp x 32 TMEM columns, W stage = grp chunks)
  int a_col0;           // first TMEM column of the A slots
  int stack;            // W chunk = [W_hi ; W_lo] stacked along N
  int acc_stride;       // TMEM columns per accumulator buffer
  int n_acc;            // accumulator buffers (2: the epilogue overlaps the next tile)
};

// ------------------------------------------------------------------------------------------------------------------
// Tile plan builder: one CTA (128 threads = tile rows) per output tile.
// ------------------------------------------------------------------------------------------------------------------
constexpr int HASH = 2048;

__device__ __forceinline__ bool hash_insert(int* keys, uint8_t* tag, int e, int k) {
  uint32_t slot = ((uint32_t)e * 2654435761u) >> 21;           // 11 bits
  while (true) {
    const int old = atomicCAS(&keys[slot], -1, e);
    if (old == -1) {
      tag[slot] = (uint8_t)k;
      return true;
    }
    if (old == e) return false;
    slot = (slot + 1) & (HASH - 1);
  }
}

__global__ void __launch_bounds__(128) tile_plan_kernel(const int* __restrict__ nbr, int koff, int m_out, int* __restrict__ hdr,
                                                        uint16_t* __restrict__ local, int* __restrict__ pool,
                                                        int* __restrict__ pool_counter) {
  __shared__ int nb[MAX_KOFF * TILE_M];
  __shared__ int keys[HASH];
  __shared__ uint8_t tag[HASH];
  __shared__ int sorted[S_CAP];
  __shared__ int warp_sums[4];
  __shared__ int s_distinct, s_base, s_any;
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int r = tile * TILE_M + tid;
  for (int k = 0; k < koff; ++k) nb[k * TILE_M + tid] = (r < m_out) ? __ldg(nbr + (size_t)k * m_out + r) : -1;
  for (int i = tid; i < HASH; i += 128) keys[i] = -1;
  if (tid == 0) s_distinct = 0;
  __syncthreads();

  int n_pass = 0;
  int* h = hdr + (size_t)tile * HDR_INTS;
  uint16_t* loc = local + (size_t)tile * koff * TILE_M;

  // close the pass over offsets [k0, k1): distinct rows inserted by those offsets -> sorted list -> pool, local table
  auto finalize = [&](int k0, int k1) {
    // compaction of the occupied slots whose first inserter is an offset of this pass
    int mine[HASH / 128];
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < HASH / 128; ++j) {
      const int s = tid * (HASH / 128) + j;
      const int e = keys[s];
      const bool in = e >= 0 && (int)tag[s] < k1;
      mine[j] = in ? e : -1;
      cnt += in ? 1 : 0;
    }
    // block exclusive scan of cnt (4 warps)
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
    __syncthreads();
    int off = incl - cnt;
    for (int w = 0; w < (tid >> 5); ++w) off += warp_sums[w];
    const int n = warp_sums[0] + warp_sums[1] + warp_sums[2] + warp_sums[3];
#pragma unroll
    for (int j = 0; j < HASH / 128; ++j)
      if (mine[j] >= 0) sorted[off++] = mine[j];
    int ns = 1;
    while (ns < n) ns <<= 1;
    __syncthreads();
    for (int i = n + tid; i < ns; i += 128) sorted[i] = 0x7fffffff;
    __syncthreads();
    // bitonic sort of sorted[0, ns)
    for (int size = 2; size <= ns; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = tid; i < (ns >> 1); i += 128) {
          const int lo = 2 * i - (i & (stride - 1));
          const int hi = lo + stride;
          const bool up = (lo & size) == 0;
          const int a = sorted[lo], b = sorted[hi];
          if ((a > b) == up) {
            sorted[lo] = b;
            sorted[hi] = a;
          }
        }
        __syncthreads();
      }
    }
    if (tid == 0) {
      s_base = n > 0 ? atomicAdd(pool_counter, (n + 3) & ~3) : 0;
      s_any = 0;
    }
    __syncthreads();
    const int base = s_base;
    for (int i = tid; i < n; i += 128) pool[base + i] = sorted[i];
    int kmask = 0;
    for (int k = k0; k < k1; ++k) {
      const int e = nb[k * TILE_M + tid];
      uint32_t pos = NONE16;
      if (e >= 0) {
        int lo = 0, hi = n - 1;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (sorted[mid] < e) lo = mid + 1;
          else hi = mid;
        }
        pos = (uint32_t)lo;
      }
      loc[k * TILE_M + tid] = (uint16_t)pos;
      if (__syncthreads_or(e >= 0)) kmask |= 1 << k;
    }
    if (tid == 0 && n_pass < MAXP) {
      h[1 + 3 * n_pass] = kmask;
      h[2 + 3 * n_pass] = base;
      h[3 + 3 * n_pass] = n;
    }
    ++n_pass;
    __syncthreads();
  };

  int k0 = 0;
  for (int k = 0; k < koff; ++k) {
    const int e = nb[k * TILE_M + tid];
    if (e >= 0 && hash_insert(keys, tag, e, k)) atomicAdd(&s_distinct, 1);
    __syncthreads();
    if (s_distinct > S_CAP) {                 // offset k does not fit any more: close [k0, k), restart the table with k
      finalize(k0, k);
      for (int i = tid; i < HASH; i += 128) keys[i] = -1;
      if (tid == 0) s_distinct = 0;
      __syncthreads();
      if (e >= 0 && hash_insert(keys, tag, e, k)) atomicAdd(&s_distinct, 1);
      __syncthreads();
      k0 = k;
    }
  }
  finalize(k0, koff);
  if (tid == 0) {
    // a tile without any pair still runs one (all-zero) step so that its accumulator is defined
    if (n_pass == 1 && h[1] == 0) h[1] = 1;
    h[0] = n_pass < MAXP ? n_pass : MAXP;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Convolution kernel
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(N_THREADS, 1) gather_gemm_once_kernel(const ls3d_gemm_args p, const Cfg cfg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t w_bytes = (uint32_t)p.n_pad * 128u;
  const uint32_t local_bytes = (uint32_t)p.koff * TILE_M * 2u;
  uint8_t* stage_s = smem;                                          // [2][S_CAP][128]
  uint8_t* w_s = stage_s + 2 * STAGE_BYTES;                         // [ts][grp][w_bytes]
  uint16_t* local_s = (uint16_t*)(w_s + cfg.ts * cfg.grp * w_bytes); // [2][koff][128]
  int* hdr_s = (int*)((uint8_t*)local_s + 2 * local_bytes);         // [2][HDR_INTS]
  uint64_t* bars = (uint64_t*)(hdr_s + 2 * HDR_INTS);
  uint32_t* tmem_slot = (uint32_t*)(bars + 4 * MAXR + 16);
  float* colv = (float*)(((uintptr_t)(tmem_slot + 4) + 15) & ~(uintptr_t)15);
  float* stg = colv + 6 * COLV;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
#ifdef LS3D_PROF
  unsigned int trc_ = 0;                        // per-thread record counter (each role is traced by one thread)
#endif
  const int cin = p.c0 + p.c1;
  const int nchunk = (p.cin_pad + KCH - 1) / KCH;
  const int ntiles = (p.m_out + TILE_M - 1) / TILE_M;

  const uint32_t full_bar0 = smem_u32(bars);                     // step operands ready [ts]: 4 splitter warps (A slot written) +
                                                                 // the W loader's arrive.expect_tx (W chunks landed)
  const uint32_t aempty_bar0 = smem_u32(bars + MAXR);            // A slot + W stage consumed [ts] (tcgen05.commit)
  const uint32_t accf_bar0 = smem_u32(bars + 3 * MAXR);          // accumulator full  [2]
  const uint32_t acce_bar0 = smem_u32(bars + 3 * MAXR + 2);      // accumulator empty [2]
  const uint32_t sfull_bar0 = smem_u32(bars + 3 * MAXR + 4);     // stage buffer filled   [2] (4 stager warps)
  const uint32_t sempty_bar0 = smem_u32(bars + 3 * MAXR + 6);    // stage buffer consumed [2] (8 splitter warps)
  const uint32_t pfull_bar0 = smem_u32(bars + 4 * MAXR);         // plan slice landed [2] (expect_tx)
  const uint32_t pempty_bar0 = smem_u32(bars + 4 * MAXR + 2);    // plan slice free   [2] (14 warps)

  constexpr uint32_t TMEM_COLS = 512;
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < MAXR; ++s) {
        mbar_init(full_bar0 + 8 * s, 5);
        mbar_init(aempty_bar0 + 8 * s, 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(accf_bar0 + 8 * b, 1);
        mbar_init(acce_bar0 + 8 * b, 4);
        mbar_init(sfull_bar0 + 8 * b, 4);
        mbar_init(sempty_bar0 + 8 * b, 8);
        mbar_init(pfull_bar0 + 8 * b, 1);
        mbar_init(pempty_bar0 + 8 * b, 14);
      }
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
  if (tid == N_THREADS - 1) TRACE(7, 0, 0);

  if (warp < SPLIT_WARP0) {
    // =========================== stagers ===========================
    // thread -> (row slot rs = tid >> 2 of a 32-row batch, quarter q = 8 channels of the 32-channel chunk)
    const int q = tid & 3;
    const int rs = tid >> 2;
    const char* const base0 = reinterpret_cast<const char*>(p.in0);
    const char* const base1 = reinterpret_cast<const char*>(p.in1);
    const size_t ldb0 = (size_t)p.ld0 * 4u, ldb1 = (size_t)p.ld1 * 4u;
    int ti = 0;
    uint32_t sg = 0;                                   // (pass, chunk) stage counter: buffer sg & 1, phase (sg >> 1) & 1
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int pb = ti & 1;
      mbar_wait(pfull_bar0 + 8 * pb, (uint32_t)(ti >> 1) & 1u);
      if (tid == 0) TRACE(0, 0, ti);
      const int* h = hdr_s + pb * HDR_INTS;
      const int n_pass = h[0];
      for (int ps = 0; ps < n_pass; ++ps) {
        const int base = h[2 + 3 * ps], cnt = h[3 + 3 * ps];
        // ids of the rows this thread stages (row rs + 32 j of the pass), kept in registers across the chunks
        int ids[S_CAP / 32];
#pragma unroll
        for (int j = 0; j < S_CAP / 32; ++j) ids[j] = (rs + 32 * j < cnt) ? __ldg(p.plan_pool + base + rs + 32 * j) : -1;
        for (int c = 0; c < nchunk; ++c, ++sg) {
          const int sb = sg & 1;
          mbar_wait(sempty_bar0 + 8 * sb, ((sg >> 1) & 1u) ^ 1u);
          if (tid == 0) TRACE(0, 1, sg);
          uint8_t* st = stage_s + sb * STAGE_BYTES;
          const int col = c * KCH + q * 8;             // first of this thread's 8 channels
          // the 8 channels never straddle in0 | in1 (c0 is a multiple of 4, handled per 16-byte half below)
#pragma unroll
          for (int j0 = 0; j0 < S_CAP / 32; j0 += 4) {
            if (j0 * 32 >= cnt) break;
            float4 v[4][2];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int id = ids[j0 + jj];
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                const int cc = col + hf * 4;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (id >= 0 && cc < cin && !(p.debug_skip & 1)) {
                  const char* src = (cc < p.c0) ? base0 + (size_t)id * ldb0 + (size_t)cc * 4
                                                : base1 + (size_t)id * ldb1 + (size_t)(cc - p.c0) * 4;
                  t = __ldg(reinterpret_cast<const float4*>(src));
                }
                v[jj][hf] = t;
              }
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int row = rs + 32 * (j0 + jj);
              if (row < cnt) {
                const float4 a = v[jj][0], b = v[jj][1];
                uint4 hi, lo;
                hi.x = pack_bf16x2(a.x, a.y); hi.y = pack_bf16x2(a.z, a.w);
                hi.z = pack_bf16x2(b.x, b.y); hi.w = pack_bf16x2(b.z, b.w);
                lo.x = pack_bf16x2(a.x - __uint_as_float(hi.x << 16), a.y - __uint_as_float(hi.x & 0xFFFF0000u));
                lo.y = pack_bf16x2(a.z - __uint_as_float(hi.y << 16), a.w - __uint_as_float(hi.y & 0xFFFF0000u));
                lo.z = pack_bf16x2(b.x - __uint_as_float(hi.z << 16), b.y - __uint_as_float(hi.z & 0xFFFF0000u));
                lo.w = pack_bf16x2(b.z - __uint_as_float(hi.w << 16), b.w - __uint_as_float(hi.w & 0xFFFF0000u));
                uint8_t* rowp = st + row * 128;
                *reinterpret_cast<uint4*>(rowp + ((q ^ (row & 7)) << 4)) = hi;
                *reinterpret_cast<uint4*>(rowp + (((4 + q) ^ (row & 7)) << 4)) = lo;
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(sfull_bar0 + 8 * sb);
          if (tid == 0) TRACE(0, 2, sg);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(pempty_bar0 + 8 * pb);
    }
  } else if (warp < EPI_WARP0) {
    // =========================== splitters ===========================
    // A STEP = up to cfg.grp active offsets of one (pass, chunk): one handshake (slot wait, full arrival, commit) per step
    // instead of per offset - the issue / synchronisation cost of a step (~550 cycles, measured with the CTA-0 timeline)
    // is what bounded the one-offset-per-step version, not the tensor pipe or the shared-memory copies.
    const int grp = (warp - SPLIT_WARP0) >> 2;         // steps with (g & 1) == grp
    const int row = ((warp & 3) << 5) | lane;          // tile row == TMEM lane
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t slot_cols = (uint32_t)(cfg.grp * A_SLOT_COLS);
    Ring tr(cfg.ts);
    int g = 0, ti = 0;
    uint32_t sg = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int pb = ti & 1;
      mbar_wait(pfull_bar0 + 8 * pb, (uint32_t)(ti >> 1) & 1u);
      const int* h = hdr_s + pb * HDR_INTS;
      const uint16_t* loc = local_s + pb * (p.koff * TILE_M) + row;
      const int n_pass = h[0];
      for (int ps = 0; ps < n_pass; ++ps) {
        const uint32_t kmask = (uint32_t)h[1 + 3 * ps];
        for (int c = 0; c < nchunk; ++c, ++sg) {
          const int sb = sg & 1;
          mbar_wait(sfull_bar0 + 8 * sb, (sg >> 1) & 1u);
          if (lane == 0 && (warp & 3) == 0) TRACE(1 + grp, 0, sg);
          const uint8_t* st = stage_s + sb * STAGE_BYTES;
          uint32_t km = kmask;
          while (km) {
            if ((g & 1) == grp) {
              if (lane == 0 && (warp & 3) == 0) TRACE(1 + grp, 1, g);
              mbar_wait(aempty_bar0 + 8 * tr.idx, tr.ph ^ 1u);
              if (lane == 0 && (warp & 3) == 0) TRACE(1 + grp, 2, g);
              tc_fence_after();
              uint32_t ta = tmem_a0 + lane_addr + (uint32_t)tr.idx * slot_cols;
              for (int j = 0; j < cfg.grp && km; ++j, ta += A_SLOT_COLS) {
                const int k = __ffs(km) - 1;
                km &= km - 1;
                if (p.debug_skip & 2) continue;
                const uint32_t s = loc[k * TILE_M];
                uint32_t hi[16], lo[16];
                if (s != NONE16) {
                  const uint8_t* rowp = st + s * 128;
                  const uint32_t x = s & 7u;
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const uint4 a = *reinterpret_cast<const uint4*>(rowp + ((u ^ x) << 4));
                    const uint4 b = *reinterpret_cast<const uint4*>(rowp + (((4 + u) ^ x) << 4));
                    hi[4 * u] = a.x; hi[4 * u + 1] = a.y; hi[4 * u + 2] = a.z; hi[4 * u + 3] = a.w;
                    lo[4 * u] = b.x; lo[4 * u + 1] = b.y; lo[4 * u + 2] = b.z; lo[4 * u + 3] = b.w;
                  }
                } else {
#pragma unroll
                  for (int u = 0; u < 16; ++u) hi[u] = lo[u] = 0u;
                }
                tmem_st16(ta, hi);
                tmem_st16(ta + 16, lo);
              }
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(full_bar0 + 8 * tr.idx);
              if (lane == 0 && (warp & 3) == 0) TRACE(1 + grp, 3, g);
            } else {
              for (int j = 0; j < cfg.grp && km; ++j) km &= km - 1;
            }
            ++g;
            tr.next();
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(sempty_bar0 + 8 * sb);    // this warp no longer reads the stage buffer
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(pempty_bar0 + 8 * pb);
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc_bf16((uint32_t)p.n_pad);
    const uint32_t idesc2 = make_idesc_bf16(2u * (uint32_t)p.n_pad);
    const uint32_t tbase = bcast0(tmem_base);
    const uint32_t ta0 = tbase + (uint32_t)cfg.a_col0;
    const uint32_t ws0 = smem_u32(w_s);
    const uint32_t slot_cols = (uint32_t)(cfg.grp * A_SLOT_COLS);
    const uint32_t wstage_bytes = (uint32_t)cfg.grp * w_bytes;
    Ring tr(cfg.ts);
    int ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int pb = ti & 1;
      const int buf = cfg.n_acc == 2 ? (ti & 1) : 0;
      mbar_wait(pfull_bar0 + 8 * pb, (uint32_t)(ti >> 1) & 1u);
      const int* h = hdr_s + pb * HDR_INTS;
      const int n_pass = (int)bcast0((uint32_t)h[0]);
      const uint32_t acc_use = cfg.n_acc == 2 ? (uint32_t)(ti >> 1) : (uint32_t)ti;
      mbar_wait(acce_bar0 + 8 * buf, (acc_use & 1u) ^ 1u);     // epilogue drained this accumulator
      if (lane == 0) TRACE(3, 3, ti);
      tc_fence_after();
      const uint32_t tacc = tbase + (uint32_t)(buf * cfg.acc_stride);
      int nst = 0;
      for (int ps = 0; ps < n_pass; ++ps)
        nst += ((__popc(bcast0((uint32_t)h[1 + 3 * ps])) + cfg.grp - 1) / cfg.grp) * nchunk;
      int st = 0;
      for (int ps = 0; ps < n_pass; ++ps) {
        const uint32_t kmask = bcast0((uint32_t)h[1 + 3 * ps]);
        for (int c = 0; c < nchunk; ++c) {
          const int nsl = min(KCH, p.cin_pad - c * KCH) / 16;
          uint32_t km = kmask;
          while (km) {
            int cnt = 0;
            for (; cnt < cfg.grp && km; ++cnt) km &= km - 1;
            mbar_wait(full_bar0 + 8 * tr.idx, tr.ph);
            if (lane == 0) TRACE(3, 0, st);
            tc_fence_after();
            const uint32_t wst = ws0 + (uint32_t)tr.idx * wstage_bytes;
            const uint32_t a_slot = ta0 + (uint32_t)tr.idx * slot_cols;
            if (elect_one()) {
              if (!(p.debug_skip & 4)) {
                for (int m = 0; m < cnt; ++m) {
                  const uint64_t bdesc = cfg.stack ? make_desc_k_sw64(wst + (uint32_t)m * w_bytes)
                                                   : make_desc_k_sw128(wst + (uint32_t)m * w_bytes);
                  const uint32_t a_hi = a_slot + (uint32_t)(m * A_SLOT_COLS);
                  const uint32_t a_lo = a_hi + 16;
                  if (cfg.stack) {
                    for (int j = 0; j < nsl; ++j) {
                      const uint64_t o = (uint64_t)(2 * j);
                      umma_bf16_ts(tacc, a_hi + 8 * j, bdesc + o, idesc2, (st > 0 || m > 0 || j > 0) ? 1u : 0u);
                      umma_bf16_ts(tacc, a_lo + 8 * j, bdesc + o, idesc, 1u);
                    }
                  } else {
                    for (int j = 0; j < nsl; ++j) {
                      const uint64_t o = (uint64_t)(2 * j);
                      umma_bf16_ts(tacc, a_hi + 8 * j, bdesc + o, idesc, (st > 0 || m > 0 || j > 0) ? 1u : 0u);
                      umma_bf16_ts(tacc, a_hi + 8 * j, bdesc + 4 + o, idesc, 1u);
                      umma_bf16_ts(tacc, a_lo + 8 * j, bdesc + o, idesc, 1u);
                    }
                  }
                }
              }
              umma_commit(aempty_bar0 + 8 * tr.idx);
              if (st == nst - 1) umma_commit(accf_bar0 + 8 * buf);
            }
            __syncwarp();
            if (lane == 0) TRACE(3, 2, st);
            ++st;
            tr.next();
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(pempty_bar0 + 8 * pb);
    }
  } else if (warp == W_WARP) {
    // =========================== W loader ===========================
    // lane m of the warp issues the copy of the step's m-th offset (bulk-copy issue costs ~500 cycles per thread)
    Ring wr(cfg.ts);
    int ti = 0;
    const uint8_t* wg = reinterpret_cast<const uint8_t*>(p.w);
    const uint32_t wstage_bytes = (uint32_t)cfg.grp * w_bytes;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int pb = ti & 1;
      mbar_wait(pfull_bar0 + 8 * pb, (uint32_t)(ti >> 1) & 1u);
      const int* h = hdr_s + pb * HDR_INTS;
      const int n_pass = (int)bcast0((uint32_t)h[0]);
      for (int ps = 0; ps < n_pass; ++ps) {
        const uint32_t kmask = bcast0((uint32_t)h[1 + 3 * ps]);
        for (int c = 0; c < nchunk; ++c) {
          uint32_t km = kmask;
          while (km) {
            int cnt = 0, myk = -1;
            for (; cnt < cfg.grp && km; ++cnt) {
              if (cnt == lane) myk = __ffs(km) - 1;
              km &= km - 1;
            }
            mbar_wait(aempty_bar0 + 8 * wr.idx, wr.ph ^ 1u);
            if (lane == 0) {
              TRACE(4, 0, cnt);
              if (p.debug_skip & 16) mbar_arrive(full_bar0 + 8 * wr.idx);
              else mbar_arrive_expect_tx(full_bar0 + 8 * wr.idx, (uint32_t)cnt * w_bytes);
            }
            __syncwarp();
            if (myk >= 0 && !(p.debug_skip & 16))
              bulk_g2s(smem_u32(w_s) + (uint32_t)wr.idx * wstage_bytes + (uint32_t)lane * w_bytes,
                       wg + ((size_t)myk * nchunk + c) * w_bytes, w_bytes, full_bar0 + 8 * wr.idx);
            wr.next();
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(pempty_bar0 + 8 * pb);
    }
  } else if (warp == PLAN_WARP) {
    // =========================== plan loader ===========================
    // one tile ahead: the tile's header (128 bytes) and local table (koff x 256 bytes) land by bulk copy
    int ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int pb = ti & 1;
      mbar_wait(pempty_bar0 + 8 * pb, ((uint32_t)(ti >> 1) & 1u) ^ 1u);
      if (lane == 0) {
        TRACE(5, 0, ti);
        mbar_arrive_expect_tx(pfull_bar0 + 8 * pb, local_bytes + HDR_INTS * 4u);
        bulk_g2s(smem_u32(hdr_s + pb * HDR_INTS), p.plan_hdr + (size_t)tile * HDR_INTS, HDR_INTS * 4u, pfull_bar0 + 8 * pb);
        bulk_g2s(smem_u32((uint8_t*)local_s + pb * local_bytes), p.plan_local + (size_t)tile * p.koff * TILE_M, local_bytes,
                 pfull_bar0 + 8 * pb);
      }
      __syncwarp();
    }
  } else {
    // =========================== epilogue ===========================
    const int q = warp & 3;
    const int et = q * 32 + lane;
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
      const int buf = cfg.n_acc == 2 ? (ti & 1) : 0;
      const uint32_t acc_use = cfg.n_acc == 2 ? (uint32_t)(ti >> 1) : (uint32_t)ti;
      mbar_wait(accf_bar0 + 8 * buf, acc_use & 1u);
      if (et == 0) TRACE(6, 0, ti);
      tc_fence_after();
      const uint32_t trow = tmem_base + (uint32_t)(buf * cfg.acc_stride) + ((uint32_t)(q * 32) << 16);
      if (!(p.debug_skip & 8)) epilogue_tile(p, trow, tile * TILE_M, et, colv, stg, cfg.stack ? (uint32_t)p.n_pad : 0u);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce_bar0 + 8 * buf);
      if (et == 0) TRACE(6, 1, ti);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tid == N_THREADS - 1) TRACE(7, 1, 0);
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

static size_t smem_bytes_for(int ts, int grp, int n_pad, int koff) {
  size_t b = 1024;  // alignment slack
  b += (size_t)2 * STAGE_BYTES + (size_t)ts * grp * n_pad * 128;
  b += (size_t)2 * koff * TILE_M * 2 + 2 * HDR_INTS * 4;
  b += (4 * MAXR + 16) * 8 + 16 + 32;
  b += (size_t)(6 * COLV + TILE_M * STG_LD) * 4;
  return b;
}

}  // namespace once
}  // namespace ls3d

#ifdef LS3D_PROF
// development: reset / read the CTA-0 event trace (8 roles x TRACE_CAP records, then the 8 counts)
extern "C" int ls3d_debug_once_trace_reset() {
  unsigned int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  return (int)cudaMemcpyToSymbol(ls3d::once::g_trace_n, z, sizeof(z));
}
extern "C" int ls3d_debug_once_trace(unsigned long long* host_out, unsigned int* counts) {
  cudaError_t e = cudaMemcpyFromSymbol(host_out, ls3d::once::g_trace, sizeof(unsigned long long) * 8 * ls3d::once::TRACE_CAP);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaMemcpyFromSymbol(counts, ls3d::once::g_trace_n, sizeof(unsigned int) * 8);
}
#endif

// ------------------------------------------------------------------------------------------------------------------ C ABI
extern "C" int ls3d_tile_plan_bytes(int32_t koff, int32_t m_out, int64_t* hdr_bytes, int64_t* local_bytes, int64_t* pool_bytes) {
  using namespace ls3d;
  if (koff < 1 || koff > MAX_KOFF || m_out < 0 || !hdr_bytes || !local_bytes || !pool_bytes) return LS3D_ERR_ARG;
  const long long tiles = ls3d_div_up(m_out, TILE_M);
  *hdr_bytes = tiles * once::HDR_INTS * 4 + 16;
  *local_bytes = tiles * koff * TILE_M * 2 + 16;
  // every staged row of every pass is backed by at least one pair of that pass (+ 3 rows of alignment padding per pass)
  *pool_bytes = ((long long)koff * tiles * TILE_M + tiles * once::MAXP * 4) * 4 + 16;
  return LS3D_OK;
}

extern "C" int ls3d_tile_plan_build(const int32_t* nbr, int32_t koff, int32_t m_out, int32_t* hdr, void* local, int32_t* pool,
                                    int32_t* pool_counter, void* stream) {
  using namespace ls3d;
  if (m_out <= 0) return LS3D_OK;
  if (!nbr || !hdr || !local || !pool || !pool_counter || koff < 1 || koff > MAX_KOFF) return LS3D_ERR_ARG;
  if (((uintptr_t)hdr & 15) || ((uintptr_t)local & 15) || ((uintptr_t)pool & 15)) return LS3D_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(pool_counter, 0, sizeof(int), st);
  once::tile_plan_kernel<<<ls3d_div_up(m_out, TILE_M), 128, 0, st>>>(nbr, koff, m_out, hdr, (uint16_t*)local, pool, pool_counter);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

// does the gather-once kernel have a shared-memory configuration for this launch?  (n_pad 256 with 27 offsets does not:
// ls3d_gather_gemm then runs the per-pair engine on the same rulebook)
int ls3d_gather_gemm_once_fits(const ls3d_gemm_args* a) {
  using namespace ls3d::once;
  return a->cin_pad % 16 == 0 && a->n_pad <= 256 && a->epi == LS3D_EPI_LINEAR &&
         smem_bytes_for(2, 1, a->n_pad, a->koff) <= 227 * 1024;
}

// called by ls3d_gather_gemm (gather_gemm.cu) for launches that carry a tile plan, after the common argument checks
int ls3d_gather_gemm_once_launch(const ls3d_gemm_args* a, int num_sms, void* stream) {
  using namespace ls3d;
  using namespace ls3d::once;
  if (a->cin_pad % 16 || a->n_pad > 256 || !a->plan_hdr || !a->plan_local || !a->plan_pool || a->epi != LS3D_EPI_LINEAR)
    return LS3D_ERR_ARG;
  if (((uintptr_t)a->plan_hdr & 15) || ((uintptr_t)a->plan_local & 15) || ((a->koff * TILE_M * 2) & 15)) return LS3D_ERR_ARG;
  Cfg cfg;
  cfg.stack = a->n_pad <= 96 ? 1 : 0;          // must match PackedWeight._pack_bf16x3 (lidarseg3d_b200/gemm.py)
  cfg.acc_stride = cfg.stack ? 2 * a->n_pad : a->n_pad;
  cfg.n_acc = 2 * cfg.acc_stride <= 384 ? 2 : 1;
  const int acc_cols = cfg.n_acc * cfg.acc_stride;
  cfg.a_col0 = acc_cols <= 256 ? 256 : 384;
  if (acc_cols > 384) return LS3D_ERR_ARG;     // n_pad 256 stacked never happens (stack needs n_pad <= 96); 256 -> 1 x 256
  // offsets per step: as many as leave >= 2 steps in flight in the TMEM columns and the shared memory that remain
  static int grp_env = -1;
  if (grp_env < 0) {
    const char* e = getenv("LS3D_ONCE_GROUP");
    grp_env = e ? atoi(e) : 0;
  }
  const int avail_cols = 512 - cfg.a_col0;
  cfg.grp = grp_env > 0 ? grp_env : 4;
  if (cfg.grp > a->koff) cfg.grp = a->koff;
  while (cfg.grp > 1 && (avail_cols / (cfg.grp * A_SLOT_COLS) < 2 || smem_bytes_for(2, cfg.grp, a->n_pad, a->koff) > 227 * 1024))
    --cfg.grp;
  int ts_max = avail_cols / (cfg.grp * A_SLOT_COLS);
  if (ts_max > MAXR) ts_max = MAXR;
  cfg.ts = ts_max;
  while (cfg.ts > 2 && smem_bytes_for(cfg.ts, cfg.grp, a->n_pad, a->koff) > 227 * 1024) --cfg.ts;
  const size_t smem = smem_bytes_for(cfg.ts, cfg.grp, a->n_pad, a->koff);
  if (smem > 227 * 1024 || cfg.ts < 1) return LS3D_ERR_ARG;
  const int ntiles = ls3d_div_up(a->m_out, TILE_M);
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  static bool optin[64] = {false};
  cudaError_t e = ls3d_optin_smem(gather_gemm_once_kernel, optin);
  if (e != cudaSuccess) return (int)e;
  gather_gemm_once_kernel<<<grid, N_THREADS, smem, (cudaStream_t)stream>>>(*a, cfg);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}
