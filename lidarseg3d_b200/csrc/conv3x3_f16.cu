// Camera stem: fused 3x3 / stride 1 / pad 1 convolution + folded BatchNorm bias (+ residual) (+ ReLU) on channels-last fp16
// maps, tcgen05 tensor cores with fp32 accumulation in TMEM, all global traffic through the TMA unit.
//
// Replaces (eval mode) the conv3x3 -> BN -> [+identity] -> ReLU pairs of the HRNet BasicBlocks, reference
// det3d/models/img_backbones/resnet_mmcv.py:20-100 as instantiated by hrnet.py:78-226 (4 blocks per branch and module; 64 of
// them per forward on the 18-channel full-resolution branch alone), which cuDNN serves with 55-130 us launches at 18-72
// channels.  fp16 operands carry the same 10-bit mantissa as the TF32 tensor-core path the fp32 stem uses.
//
// Implicit GEMM without im2col traffic: a persistent CTA walks 16 x 8 output tiles (128 pixels = the UMMA M dimension).
// The (16+2) x (8+2) input halo of a tile is staged ONCE in shared memory, channel-chunk major ([16-byte chunk][pixel]) -
// the un-swizzled K-major UMMA layout whose 8-row groups are 8 x-adjacent pixels - so each of the 9 taps is the same buffer
// read through a shared-memory descriptor whose start address is shifted by (dy * 10 + dx) pixels.  The K dimension of the
// GEMM is the list of (tap, 8-channel chunk) pairs; one MMA (K = 16) takes two consecutive list entries wherever they lie in
// the halo buffer (the descriptor's leading-dimension byte offset is per MMA), so 18 padded channels cost ceil(27 / 2) = 14
// MMAs per tile rather than 9 * 2.  Weights are packed on the device in that order and stay resident in shared memory.
//   warp 0      producer: one thread; per tile one 4-D tensor-map copy per channel chunk (box 8 ch x 10 x 18, zero fill outside
//                         the image = the conv padding) + one for the residual tile, completion on the stage's mbarrier
//   warp 1      MMA     : one elected thread, tcgen05.mma.kind::f16 M=128 N=Cout K=16, four accumulators in tensor memory
//   warps 2-9   epilogue: two teams of 4 warps (even / odd tiles of the CTA): tcgen05.ld -> + bias + residual
//                         (read from the stage) -> ReLU -> fp16, written in place over the residual tile, then ONE tensor-map
//                         store per tile (clips partial tiles); overlaps the following tiles' MMAs
// Bound: HBM (one read of the input, one of the residual, one write) / the tensor core's shared-memory operand fetch
// (128 rows x 32 bytes per MMA).
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "../../include/ls3d.h"

namespace ls3d {
namespace c3 {

constexpr int TW = 8, TH = 16;              // output tile: 8 wide x 16 tall
constexpr int HALO_W = TW + 2, HALO_H = TH + 2;
constexpr int HPIX = HALO_W * HALO_H;       // 180 halo pixels
constexpr int CH_STRIDE = 2944;             // bytes between channel chunks of a 3x3 halo buffer (180 * 16 rounded up to 128)
constexpr int CH_STRIDE_1X1 = 2048;         // 1x1: the staged box is the 128-pixel tile itself
__host__ __device__ __forceinline__ int ch_stride_of(int ntap) { return ntap == 1 ? CH_STRIDE_1X1 : CH_STRIDE; }
constexpr int PROD_WARP = 0, MMA_WARP = 1, EPI_WARP0 = 2, N_EPI_WARPS = 8;   // two epilogue teams of 4 warps alternate tiles
constexpr int N_THREADS = (EPI_WARP0 + N_EPI_WARPS) * 32;
constexpr int MAX_BUF = 8;
constexpr int MAX_MMA = 144;                 // ceil(9 * 32 / 2): cin <= 256

constexpr int MAX_PASS = 24;
constexpr int PASS_BIAS = 1, PASS_RES_EXT = 2, PASS_RES_OUT = 4, PASS_RELU = 8;

struct Args {
  const __half* w;        // packed weights, see pack kernel (pass 0; pass i: pass_w[i])
  const float* bias;      // fp32 bias of the whole output tensor (indexed from out_c_off) or NULL
  int has_res;            // single-pass launches: residual present (multi-pass: per-pass flags)
  // A launch may run several PASSES over the same tiles: each pass has its own resident weight block, input-channel and
  // output-channel slice; passes over the input slices of one output slice accumulate through the fp32 output map (a CTA
  // always owns the same tiles, so pass i + 1 reads what the same CTA wrote in pass i).  One launch per convolution instead
  // of one per slice: no launch / tail / tensor-memory set-up cost between the slices.
  int n_pass, use_table;
  const __half* pass_w[MAX_PASS];
  short pass_in_off[MAX_PASS], pass_out_off[MAX_PASS];
  unsigned char pass_flags[MAX_PASS];
  int dual;               // 1: fp32 residual / output maps + an fp16 operand copy of the output (see ls3d_conv_f16_dual)
  int nb;                 // rows of the B operand = accumulator columns per tile: n_pad, or 2 n_pad with split weights
                          // ([W_hi ; W_lo] stacked along N: exact fp32 weights at the cost of a wider MMA, the epilogue adds
                          // the two halves)
  int n_img, H, W, cin, cout, kcg, kc, n_mma, n_pad, relu, tiles_x, tiles_y, n_tiles, nbuf;   // H, W: OUTPUT map size
  int nphase;             // 1: stride 1; 4: stride 2 - the halo holds the four (row parity, column parity) phase planes of the
                          // input, each loaded through its own tensor map (same memory, doubled pixel strides)
  int kdata;              // nphase * kcg data chunks per halo buffer (chunk kdata = the all-zero chunk when kc > kdata)
  int in_c_off, out_c_off;   // channel offsets of this launch's slices inside the input / output (and residual) tensors
  int ch_stride;             // bytes between the channel chunks of a staged box
  int halo_w, halo_pix, org; // staged input box: (TW + 2) x (TH + 2) from (-1, -1) for 3x3; exactly the TW x TH tile for 1x1
  float inv_tiles_x, inv_tiles_per_img;
  // per MMA: low word of the A descriptor less the stage base = (offset of the lower K entry >> 4) | (LBO >> 4) << 16.
  // Lives in the kernel parameters so that the issue loop reads it with a warp-uniform constant load: descriptors fetched
  // from shared memory by the issuing thread are not provably uniform and ptxas then wraps every MMA in an
  // elect / R2UR / branch sequence.
  uint32_t a_lo[MAX_MMA];
};

// K-major, no swizzle: core matrix = 8 rows x 16 bytes, rows 16 bytes apart; `sbo` between 8-row groups, `lbo` between the
// two 16-byte K chunks of one MMA (cute::UMMA canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units)
__device__ __forceinline__ uint64_t make_desc_k_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_f16(uint32_t n) {
  uint32_t d = 0;
  d |= 1u << 4;            // c_format = F32; a_format = b_format = F16 (0); both K-major
  d |= (n >> 3) << 17;
  d |= (128u >> 4) << 24;
  return d;
}

// ntap = 9 (3x3) or 1 (1x1: the same kernel reading only the centre of the halo).
// Entry c of the K list: tap = c / kcg, channel chunk = c % kcg; entries past ntap * kcg are the all-zero chunk (index kcg of the
// halo buffer, never written by the copies).  Byte offset of its first row inside a halo buffer:
// stride 2 (nphase = 4): tap row dy reads input row 2 oy + dy - 1 = phase plane (dy odd ? even rows : odd rows) at block row
// oy - 1 (dy = 0) or oy (dy = 1, 2); the halo origin is block (ox0 - 1, oy0 - 1), so the in-halo offset is 0 or 1.  Columns alike.
__host__ __device__ __forceinline__ int k_entry_offset(int c, int kcg, int ntap, int nphase = 1) {
  if (c >= ntap * kcg) return nphase * kcg * ch_stride_of(ntap);
  const int t = c / kcg, kc = c - t * kcg;
  if (ntap == 1) return kc * CH_STRIDE_1X1;              // 1x1: the staged box is the tile itself
  const int tap = t;
  const int dy = tap / 3, dx = tap % 3;
  if (nphase == 1) return kc * CH_STRIDE + (dy * HALO_W + dx) * 16;
  const int by = dy == 0 ? 0 : 1, py = dy == 1 ? 0 : 1;
  const int bx = dx == 0 ? 0 : 1, px = dx == 1 ? 0 : 1;
  return ((py * 2 + px) * kcg + kc) * CH_STRIDE + (by * HALO_W + bx) * 16;
}
// MMA j multiplies K-list entries 2j and 2j+1; the one at the lower shared-memory offset is the first 8 of its 16 K values.
__host__ __device__ __forceinline__ void mma_entries(int j, int kcg, int ntap, int nphase, int& first, int& second) {
  const int a = 2 * j, b = 2 * j + 1;
  if (k_entry_offset(a, kcg, ntap, nphase) <= k_entry_offset(b, kcg, ntap, nphase)) {
    first = a;
    second = b;
  } else {
    first = b;
    second = a;
  }
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_team(int team) { asm volatile("bar.sync %0, 128;" ::"r"(3 + team) : "memory"); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}

struct InMaps {
  CUtensorMap m[4];
};

struct TileXY {
  int img, ty, tx;
};
// tile -> (image, tile row, tile column); float quotient + one-step fix-up, exact for tile < 2^22
__device__ __forceinline__ TileXY tile_coords(int tile, const Args& p) {
  const int tpi = p.tiles_x * p.tiles_y;
  int img = (int)((float)tile * p.inv_tiles_per_img);
  int rem = tile - img * tpi;
  if (rem < 0) {
    --img;
    rem += tpi;
  } else if (rem >= tpi) {
    ++img;
    rem -= tpi;
  }
  int ty = (int)((float)rem * p.inv_tiles_x);
  int tx = rem - ty * p.tiles_x;
  if (tx < 0) {
    --ty;
    tx += p.tiles_x;
  } else if (tx >= p.tiles_x) {
    ++ty;
    tx -= p.tiles_x;
  }
  return {img, ty, tx};
}

// Epilogue arithmetic of one output tile for thread m (= tile pixel = accumulator row): TMEM accumulator (+ the W_lo partial
// sums at column lo_col) + bias + residual (read from the staged io tile) -> ReLU -> io tile in place (dual: fp32 tile + fp16
// operand copy behind it; else fp16 tile).  `tile_io` = start of the stage's io area.
__device__ __forceinline__ void epilogue_rows(const Args& p, uint8_t* tile_io, uint32_t io32_bytes, int m, uint32_t trow,
                                              uint32_t lo_col, bool has_res, bool relu, const float* bias_s) {
  const int n_groups = p.n_pad >> 4;
  uint8_t* io = tile_io + (size_t)m * p.cout * 2;
  if (p.dual) {
    // fp32 residual stream: acc + bias + residual (fp32, read from the stage) -> ReLU -> fp32 in place + fp16 operand copy
    float* io32 = reinterpret_cast<float*>(tile_io) + (size_t)m * p.cout;
    __half* io16 = reinterpret_cast<__half*>(tile_io + io32_bytes) + (size_t)m * p.cout;
    for (int g = 0; g < n_groups; ++g) {
      const int c0 = g * 16;
      uint32_t raw[16], raw2[16];
      tmem_ld16(trow + c0, raw);
      if (lo_col) tmem_ld16(trow + lo_col + c0, raw2);
      float4 rr[4];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4)
        rr[q4] = (has_res && c0 + 4 * q4 < p.cout) ? *reinterpret_cast<const float4*>(io32 + c0 + 4 * q4)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      tmem_ld_wait();
      if (lo_col) {
#pragma unroll
        for (int e = 0; e < 16; ++e) raw[e] = __float_as_uint(__uint_as_float(raw[e]) + __uint_as_float(raw2[e]));
      }
#pragma unroll
      for (int g8 = 0; g8 < 2; ++g8) {
        if (c0 + 8 * g8 >= p.cout) continue;
        float v[8];
#pragma unroll
        for (int q4 = 0; q4 < 2; ++q4) {
          const float4 bb = *reinterpret_cast<const float4*>(bias_s + c0 + 8 * g8 + 4 * q4);
          const float4 r4 = rr[2 * g8 + q4];
          v[4 * q4 + 0] = __uint_as_float(raw[8 * g8 + 4 * q4 + 0]) + bb.x + r4.x;
          v[4 * q4 + 1] = __uint_as_float(raw[8 * g8 + 4 * q4 + 1]) + bb.y + r4.y;
          v[4 * q4 + 2] = __uint_as_float(raw[8 * g8 + 4 * q4 + 2]) + bb.z + r4.z;
          v[4 * q4 + 3] = __uint_as_float(raw[8 * g8 + 4 * q4 + 3]) + bb.w + r4.w;
        }
        if (relu) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
        }
        *reinterpret_cast<float4*>(io32 + c0 + 8 * g8) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(io32 + c0 + 8 * g8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        uint4 o;
        __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) o2[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
        *reinterpret_cast<uint4*>(io16 + c0 + 8 * g8) = o;
      }
    }
  } else
  for (int g = 0; g < n_groups; ++g) {
    const int c0 = g * 16;
    uint32_t raw[16], raw2[16];
    tmem_ld16(trow + c0, raw);
    if (lo_col) tmem_ld16(trow + lo_col + c0, raw2);
    uint4 rz[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    if (has_res) {
#pragma unroll
      for (int g8 = 0; g8 < 2; ++g8)
        if (c0 + 8 * g8 < p.cout) rz[g8] = *reinterpret_cast<const uint4*>(io + (c0 + 8 * g8) * 2);
    }
    tmem_ld_wait();
    if (lo_col) {
#pragma unroll
      for (int e = 0; e < 16; ++e) raw[e] = __float_as_uint(__uint_as_float(raw[e]) + __uint_as_float(raw2[e]));
    }
#pragma unroll
    for (int g8 = 0; g8 < 2; ++g8) {
      if (c0 + 8 * g8 >= p.cout) continue;
      const __half2* r2 = reinterpret_cast<const __half2*>(&rz[g8]);
      const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c0 + 8 * g8);
      const float4 b1 = *reinterpret_cast<const float4*>(bias_s + c0 + 8 * g8 + 4);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint4 o;
      __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 rr = __half22float2(r2[e]);
        float v0 = __uint_as_float(raw[8 * g8 + 2 * e]) + bb[2 * e] + rr.x;
        float v1 = __uint_as_float(raw[8 * g8 + 2 * e + 1]) + bb[2 * e + 1] + rr.y;
        if (relu) {
          v0 = fmaxf(v0, 0.f);
          v1 = fmaxf(v1, 0.f);
        }
        o2[e] = __floats2half2_rn(v0, v1);
      }
      *reinterpret_cast<uint4*>(io + (c0 + 8 * g8) * 2) = o;
    }
  }
}

__global__ void __launch_bounds__(N_THREADS, 1)
    conv3x3_f16_kernel(const Args p, const __grid_constant__ InMaps maps_in, const __grid_constant__ CUtensorMap map_res,
                       const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_out16) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t w_bytes = (uint32_t)p.n_mma * 2u * p.nb * 16u;
  const uint32_t halo_bytes = (uint32_t)p.kc * (uint32_t)p.ch_stride;
  const uint32_t io32_bytes = p.dual ? 128u * p.cout * 4u : 0u;      // dual: fp32 residual / output tile, then the fp16 copy
  const uint32_t io_bytes = io32_bytes + 128u * p.cout * 2u;
  const uint32_t stage_bytes = halo_bytes + io_bytes;                 // both multiples of 128
  uint8_t* stage_s = smem;                                             // [nbuf][halo | io tile]
  uint8_t* w_s = stage_s + (size_t)p.nbuf * stage_bytes;
  float* bias_s = (float*)(w_s + w_bytes);
  uint64_t* bars = (uint64_t*)(bias_s + p.n_pad);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * MAX_BUF + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t full_bar0 = smem_u32(bars);                   // stage landed  [nbuf] (1 arrival + tx bytes)
  const uint32_t empty_bar0 = smem_u32(bars + MAX_BUF);        // stage free    [nbuf] (tcgen05.commit + store drained)
  const uint32_t accf_bar0 = smem_u32(bars + 2 * MAX_BUF);     // accumulator full  [nacc <= 4]
  const uint32_t acce_bar0 = smem_u32(bars + 2 * MAX_BUF + 4); // accumulator empty [nacc <= 4]
  const int nacc = p.nb <= 128 ? 4 : 2;                        // accumulators in tensor memory (nacc * nb <= 512 columns)
  const uint32_t lo_col = p.nb > p.n_pad ? (uint32_t)p.n_pad : 0u;   // split weights: column offset of the W_lo partial sums

  // ---- one-time staging: the all-zero K chunk of every halo buffer (weights / bias: per pass, below)
  if (p.kdata < p.kc) {
    const int per = p.ch_stride / 16;
    for (int i = tid; i < p.nbuf * per; i += N_THREADS)
      reinterpret_cast<uint4*>(stage_s + (size_t)(i / per) * stage_bytes + (size_t)p.kdata * p.ch_stride)[i % per] =
          make_uint4(0, 0, 0, 0);
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(nacc * p.nb)) tmem_cols <<= 1;
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < MAX_BUF; ++s) {
        mbar_init(full_bar0 + 8 * s, 1);
        mbar_init(empty_bar0 + 8 * s, 2);
      }
      for (int b = 0; b < 4; ++b) {
        mbar_init(accf_bar0 + 8 * b, 1);
        mbar_init(acce_bar0 + 8 * b, 4);             // one arrival per epilogue warp of the team
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  } else if (warp == PROD_WARP && lane == 0) {
    for (int ph = 0; ph < p.nphase; ++ph) prefetch_map(&maps_in.m[ph]);
    prefetch_map(&map_res);
    prefetch_map(&map_out);
    if (p.dual) prefetch_map(&map_out16);
  }
  fence_proxy_async_smem();          // the zero chunk is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t stage0 = smem_u32(stage_s);

  // pipeline state of each role, carried across the passes
  int st_b = 0, st_it = 0;                    // producer / MMA: ring slot, tile counter
  uint32_t st_ph = warp == PROD_WARP ? 1u : 0u;
  const int e_team = (warp - EPI_WARP0) >> 2;
  int e_b = 0, e_bprev = 0, e_it = 0;         // epilogue team state
  uint32_t e_ph = 0;
  bool e_first = true;
  if (warp >= EPI_WARP0) {
    e_b = e_team % p.nbuf;
    e_it = e_team;
    e_ph = (uint32_t)(e_team / p.nbuf) & 1u;
  }
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA per pass

  for (int pass = 0; pass < p.n_pass; ++pass) {
  const int pflags = p.use_table ? (int)p.pass_flags[pass] : (PASS_BIAS | (p.has_res ? PASS_RES_EXT : 0) | (p.relu ? PASS_RELU : 0));
  const int in_c_off = p.use_table ? (int)p.pass_in_off[pass] : p.in_c_off;
  const int out_c_off = p.use_table ? (int)p.pass_out_off[pass] : p.out_c_off;
  const bool has_res = (pflags & (PASS_RES_EXT | PASS_RES_OUT)) != 0;
  const bool relu = (pflags & PASS_RELU) != 0;
  {
    // this pass's weights (resident) and bias; the previous pass's MMAs and stores are complete (barrier at the loop end)
    const uint4* src = reinterpret_cast<const uint4*>(p.use_table ? p.pass_w[pass] : p.w);
    uint4* dst = reinterpret_cast<uint4*>(w_s);
    for (uint32_t i = tid; i < w_bytes / 16; i += N_THREADS) dst[i] = __ldg(src + i);
    for (int c = tid; c < p.n_pad; c += N_THREADS)
      bias_s[c] = (p.bias && (pflags & PASS_BIAS) && c < p.cout) ? __ldg(p.bias + (p.use_table ? out_c_off : 0) + c) : 0.f;
    fence_proxy_async_smem();
    __syncthreads();
  }

  if (warp == PROD_WARP) {
    // =========================== producer (tensor-map copies) ===========================
    if (lane == 0) {
      if (pass > 0) asm volatile("fence.proxy.async;" ::: "memory");
      const uint32_t tx_bytes = (uint32_t)p.kdata * ((uint32_t)p.halo_pix * 16u) + (has_res ? (p.dual ? io32_bytes : io_bytes) : 0u);
      int b = st_b;
      uint32_t ph = st_ph;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        mbar_wait(empty_bar0 + 8 * b, ph);
        const TileXY t = tile_coords(tile, p);
        const uint32_t dst = stage0 + (uint32_t)b * stage_bytes;
        const uint32_t bar = full_bar0 + 8 * b;
        mbar_arrive_expect_tx(bar, tx_bytes);
        for (int ph = 0; ph < p.nphase; ++ph)
          for (int kc = 0; kc < p.kcg; ++kc)
            tma_load_4d(dst + (uint32_t)((ph * p.kcg + kc) * p.ch_stride), &maps_in.m[ph], bar, in_c_off + kc * 8, t.tx * TW - p.org,
                        t.ty * TH - p.org, t.img);
        if (has_res)
          tma_load_4d(dst + halo_bytes, (pflags & PASS_RES_OUT) ? &map_out : &map_res, bar, out_c_off, t.tx * TW, t.ty * TH, t.img);
        if (++b == p.nbuf) {
          b = 0;
          ph ^= 1u;
        }
      }
      st_b = b;
      st_ph = ph;
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc_f16((uint32_t)p.nb);
    const uint32_t tbase = bcast0(tmem_base);
    const uint32_t a_hi = (((uint32_t)p.halo_w * 16u) >> 4) | (1u << 14);   // SBO (8-row group pitch) | descriptor version
    const uint32_t b_hi = (128u >> 4) | (1u << 14);
    const uint32_t b_lo0 = (smem_u32(w_s) >> 4) | ((((uint32_t)p.nb * 16u) >> 4) << 16);
    const uint32_t b_step = 2u * (uint32_t)p.nb;                    // one MMA's weights: [2][nb][16 B]
    int b = st_b, it = st_it;
    uint32_t ph = st_ph;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int ab = it & (nacc - 1);
      mbar_wait(full_bar0 + 8 * b, ph);
      mbar_wait(acce_bar0 + 8 * ab, ((uint32_t)(it / nacc) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t tacc = tbase + (uint32_t)(ab * p.nb);
      // all operands warp-uniform: kernel parameters, loop counters, lane-0 broadcasts (start-address field: no carry out
      // of its 14 bits)
      const uint32_t a_base = (stage0 + (uint32_t)b * stage_bytes) >> 4;
      if (elect_one()) {
        for (int j = 0; j < p.n_mma; ++j) {
          const uint64_t adesc = ((uint64_t)a_hi << 32) | (uint64_t)(p.a_lo[j] + a_base);
          const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo0 + (uint32_t)j * b_step);
          umma_bf16_ss(tacc, adesc, bdesc, idesc, j > 0 ? 1u : 0u);
        }
        umma_commit(empty_bar0 + 8 * b);
        umma_commit(accf_bar0 + 8 * ab);
      }
      __syncwarp();
      if (++b == p.nbuf) {
        b = 0;
        ph ^= 1u;
      }
    }
    st_b = b;
    st_it = it;
    st_ph = ph;
  } else {
    // =========================== epilogue ===========================
    const int ew = warp - EPI_WARP0;
    const int q = warp & 3;                       // TMEM lane quarter this warp may read (warp id mod 4)
    const int team = ew >> 2;                     // team 0: even tiles of this CTA (accumulator 0), team 1: odd tiles
    const int m = q * 32 + lane;                  // tile pixel = accumulator row
    const bool leader = ((ew & 3) == 0 && lane == 0);
    const bool deep = p.nbuf >= 4;                // a stage may stay occupied until this team's next tile
    // this team takes the CTA's tiles whose running index (over all passes) has its parity
    int b = e_b, b_prev = e_bprev, it = e_it;
    uint32_t ph = e_ph;
    bool first = e_first;
    const int it_begin = pass * my_tiles, it_end = it_begin + my_tiles;
    for (; it < it_end; it += 2) {
      const int tile = blockIdx.x + (it - it_begin) * gridDim.x;
      const int ab = it & (nacc - 1);
      const uint32_t trow = tmem_base + (uint32_t)(ab * p.nb) + ((uint32_t)(q * 32) << 16);
      if (has_res) mbar_wait(full_bar0 + 8 * b, ph);
      mbar_wait(accf_bar0 + 8 * ab, (uint32_t)(it / nacc) & 1u);
      tc_fence_after();
      epilogue_rows(p, stage_s + (size_t)b * stage_bytes + halo_bytes, io32_bytes, m, trow, lo_col, has_res, relu, bias_s);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce_bar0 + 8 * ab);
      fence_proxy_async_smem();                   // this thread's output row -> visible to the bulk store (async proxy)
      bar_sync_team(team);
      if (leader) {
        const TileXY t = tile_coords(tile, p);
        tma_store_4d(&map_out, stage0 + (uint32_t)b * stage_bytes + halo_bytes, out_c_off, t.tx * TW, t.ty * TH, t.img);
        if (p.dual)
          tma_store_4d(&map_out16, stage0 + (uint32_t)b * stage_bytes + halo_bytes + io32_bytes, out_c_off, t.tx * TW, t.ty * TH,
                       t.img);
        bulk_commit();
        if (deep) {
          if (!first) {
            bulk_wait_read<1>();                  // this team's previous store has drained its stage
            mbar_arrive(empty_bar0 + 8 * b_prev);
          }
        } else {
          bulk_wait_read<0>();
          mbar_arrive(empty_bar0 + 8 * b);
        }
      }
      first = false;
      b_prev = b;
      b += 2;                                     // this team's next tile is two ring slots on
      while (b >= p.nbuf) {
        b -= p.nbuf;
        ph ^= 1u;
      }
    }
    if (leader) {
      // the pass's last store of this team has drained its stage and is visible before the next pass re-reads / re-stages
      bulk_wait_all();
      if (deep && !first) mbar_arrive(empty_bar0 + 8 * b_prev);
    }
    first = true;                                 // the deferred stage release above closed this pass
    e_b = b;
    e_bprev = b_prev;
    e_it = it;
    e_ph = ph;
    e_first = first;
  }
  tc_fence_before();
  __syncthreads();                                // end of the pass: every MMA consumed, every store complete
  tc_fence_after();
  }  // pass loop
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, tmem_cols);
}



// Epilogue of one tile pixel WITHOUT a staged io tile (K-block kernel: the maps of its convolutions are small, so the shared
// memory goes to staged halos and the weight ring instead): the pixel's residual row is fetched from global memory before the
// accumulator is awaited (whole row in registers), results go out with 16-byte global stores.
constexpr int DIRECT_MAX_C = 128;
struct DirectRes {
  float4 v[DIRECT_MAX_C / 4];
};
__device__ __forceinline__ void direct_prefetch(const Args& p, DirectRes& r, const void* res, size_t pix, int ct, int c_off,
                                                bool live) {
  if (p.dual) {
    const float* src = reinterpret_cast<const float*>(res) + pix * ct + c_off;
#pragma unroll
    for (int i = 0; i < DIRECT_MAX_C / 4; ++i)
      r.v[i] = (live && 4 * i < p.cout) ? __ldg(reinterpret_cast<const float4*>(src) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    const __half* src = reinterpret_cast<const __half*>(res) + pix * ct + c_off;
#pragma unroll
    for (int i = 0; i < DIRECT_MAX_C / 8; ++i) {
      uint4 t = (live && 8 * i < p.cout) ? __ldg(reinterpret_cast<const uint4*>(src) + i) : make_uint4(0, 0, 0, 0);
      const __half2* h2 = reinterpret_cast<const __half2*>(&t);
      const float2 a = __half22float2(h2[0]), b = __half22float2(h2[1]), c = __half22float2(h2[2]), d = __half22float2(h2[3]);
      r.v[2 * i] = make_float4(a.x, a.y, b.x, b.y);
      r.v[2 * i + 1] = make_float4(c.x, c.y, d.x, d.y);
    }
  }
}
__device__ __forceinline__ void epilogue_rows_direct(const Args& p, const DirectRes& r, bool has_res, float* out32, __half* out16,
                                                     size_t pix, int ct, int c_off, bool live, uint32_t trow, uint32_t lo_col,
                                                     bool relu, const float* bias_s) {
  const int n_groups = p.n_pad >> 4;
  float* o32 = out32 ? out32 + pix * ct + c_off : nullptr;
  __half* o16 = out16 + pix * ct + c_off;
#pragma unroll
  for (int g = 0; g < DIRECT_MAX_C / 16; ++g) {
    if (g >= n_groups) break;
    const int c0 = g * 16;
    uint32_t raw[16], raw2[16];
    tmem_ld16(trow + c0, raw);
    if (lo_col) tmem_ld16(trow + lo_col + c0, raw2);
    tmem_ld_wait();
    if (lo_col) {
#pragma unroll
      for (int e = 0; e < 16; ++e) raw[e] = __float_as_uint(__uint_as_float(raw[e]) + __uint_as_float(raw2[e]));
    }
#pragma unroll
    for (int g8 = 0; g8 < 2; ++g8) {
      if (c0 + 8 * g8 >= p.cout) continue;
      float v[8];
#pragma unroll
      for (int q4 = 0; q4 < 2; ++q4) {
        const float4 bb = *reinterpret_cast<const float4*>(bias_s + c0 + 8 * g8 + 4 * q4);
        const float4 r4 = has_res ? r.v[4 * g + 2 * g8 + q4] : make_float4(0.f, 0.f, 0.f, 0.f);
        v[4 * q4 + 0] = __uint_as_float(raw[8 * g8 + 4 * q4 + 0]) + bb.x + r4.x;
        v[4 * q4 + 1] = __uint_as_float(raw[8 * g8 + 4 * q4 + 1]) + bb.y + r4.y;
        v[4 * q4 + 2] = __uint_as_float(raw[8 * g8 + 4 * q4 + 2]) + bb.z + r4.z;
        v[4 * q4 + 3] = __uint_as_float(raw[8 * g8 + 4 * q4 + 3]) + bb.w + r4.w;
      }
      if (relu) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
      }
      if (live) {
        if (o32) {
          *reinterpret_cast<float4*>(o32 + c0 + 8 * g8) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(o32 + c0 + 8 * g8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        uint4 o;
        __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) o2[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
        *reinterpret_cast<uint4*>(o16 + c0 + 8 * g8) = o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K-block variant for the WEIGHT-HEAVY, PIXEL-LIGHT convolutions (the 72- and 144-channel HRNet branches at 40x60 / 20x30:
// 85-340 output tiles against 190-750 KB of split weights).  The resident-weight kernel above has to cut such a convolution
// into 3-12 output / input channel slices that each fit shared memory, and every slice pays the tensor core's per-MMA
// A-operand fetch again at a small N.  Here the weights are NOT resident: a work item is (group of T tiles, output slice);
// the T halos of the group stay staged in shared memory with T accumulators in tensor memory, and the slice's weight image
// streams ONCE per item through a two-slot ring (cp.async.bulk of G consecutive MMAs' B tiles, which the packed layout
// already stores contiguously) while the MMA warp walks block-outer / tile-inner: every MMA runs at the slice's full N and
// a weight byte is fetched once per T tiles.  3x3 with stride 1 or 2 (phase-plane halos as in the kernel above; 1x1 works too
// but loses to the staged TMA stores at high resolution); epilogue = epilogue_rows_direct (same arithmetic, global loads / stores).
constexpr int KB_WSLOTS = 3;               // weight ring depth: bytes in flight from L2 while the MMAs of a block run
struct KbArgs {
  int T;                      // tiles per group = stages = accumulators (T * nb <= 512 TMEM columns)
  int n_groups;               // ceil(n_tiles / T)
  int n_os;                   // output slices (work item = group * n_os + slice)
  int g_mma;                  // MMAs per weight block
  int n_blocks;               // ceil(n_mma / g_mma)
  uint32_t wslot_bytes;       // g_mma * 2 * nb * 16
  const void* res;            // residual map (fp32 in dual launches, else fp16) or NULL; out32 (dual) / out16 maps; their
  float* out32;               //   channel count out_ct
  __half* out16;
  int out_ct;
  const __half* os_w[8];      // per output slice: packed weights, output channel offset, flags (PASS_*)
  short os_out_off[8];
  unsigned char os_flags[8];
};

__global__ void __launch_bounds__(N_THREADS, 1)
    conv3x3_kb_kernel(const Args p, const KbArgs kb, const __grid_constant__ InMaps maps_in, const __grid_constant__ CUtensorMap map_res,
                      const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_out16) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t stage_bytes = (uint32_t)p.kc * (uint32_t)p.ch_stride;       // a stage is the halo alone (direct epilogue)
  uint8_t* stage_s = smem;                                             // [T][halo]
  uint8_t* w_s = stage_s + (size_t)kb.T * stage_bytes;                 // [KB_WSLOTS][wslot_bytes]
  float* bias_s = (float*)(w_s + KB_WSLOTS * (size_t)kb.wslot_bytes);  // [n_os][n_pad]
  uint64_t* bars = (uint64_t*)(bias_s + kb.n_os * p.n_pad);
  uint32_t* tmem_slot = (uint32_t*)(bars + 4 * MAX_BUF + 2 * KB_WSLOTS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t full_bar0 = smem_u32(bars);                   // halo of stage t landed   [T]
  const uint32_t empty_bar0 = smem_u32(bars + MAX_BUF);        // stage t free: the item's MMAs retired [T] (tcgen05.commit)
  const uint32_t accf_bar0 = smem_u32(bars + 2 * MAX_BUF);     // accumulator t complete [T]
  const uint32_t acce_bar0 = smem_u32(bars + 3 * MAX_BUF);     // accumulator t drained  [T] (4 epilogue warps)
  const uint32_t wfull_bar0 = smem_u32(bars + 4 * MAX_BUF);    // weight slot landed   [KB_WSLOTS]
  const uint32_t wempty_bar0 = smem_u32(bars + 4 * MAX_BUF + KB_WSLOTS);   // weight slot consumed (tcgen05.commit)
  const uint32_t lo_col = p.nb > p.n_pad ? (uint32_t)p.n_pad : 0u;

  if (p.kdata < p.kc) {
    const int per = p.ch_stride / 16;
    for (int i = tid; i < kb.T * per; i += N_THREADS)
      reinterpret_cast<uint4*>(stage_s + (size_t)(i / per) * stage_bytes + (size_t)p.kdata * p.ch_stride)[i % per] =
          make_uint4(0, 0, 0, 0);
  }
  for (int i = tid; i < kb.n_os * p.n_pad; i += N_THREADS) {
    const int os = i / p.n_pad, c = i % p.n_pad;
    bias_s[i] = (p.bias && (kb.os_flags[os] & PASS_BIAS) && c < p.cout) ? __ldg(p.bias + kb.os_out_off[os] + c) : 0.f;
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(kb.T * p.nb)) tmem_cols <<= 1;
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < MAX_BUF; ++s) {
        mbar_init(full_bar0 + 8 * s, 1);
        mbar_init(empty_bar0 + 8 * s, 1);
        mbar_init(accf_bar0 + 8 * s, 1);
        mbar_init(acce_bar0 + 8 * s, 4);
      }
      for (int s = 0; s < KB_WSLOTS; ++s) {
        mbar_init(wfull_bar0 + 8 * s, 1);
        mbar_init(wempty_bar0 + 8 * s, 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  } else if (warp == PROD_WARP && lane == 0) {
    for (int ph = 0; ph < p.nphase; ++ph) prefetch_map(&maps_in.m[ph]);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t stage0 = smem_u32(stage_s);
  const int n_items = kb.n_groups * kb.n_os;

  if (warp == PROD_WARP) {
    // =========================== producer: halos (+ residual tiles) of the item, then its weight blocks ===========================
    if (lane == 0) {
      uint32_t it = 0, wc = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int group = item / kb.n_os, os = item - group * kb.n_os;
        const uint32_t tx_bytes = (uint32_t)p.kdata * ((uint32_t)p.halo_pix * 16u);
        for (int t = 0; t < kb.T; ++t) {
          const int tile = group * kb.T + t;
          if (tile >= p.n_tiles) break;
          mbar_wait(empty_bar0 + 8 * t, (it & 1u) ^ 1u);
          const TileXY xy = tile_coords(tile, p);
          const uint32_t dst = stage0 + (uint32_t)t * stage_bytes;
          const uint32_t bar = full_bar0 + 8 * t;
          mbar_arrive_expect_tx(bar, tx_bytes);
          for (int ph = 0; ph < p.nphase; ++ph)
            for (int kc = 0; kc < p.kcg; ++kc)
              tma_load_4d(dst + (uint32_t)((ph * p.kcg + kc) * p.ch_stride), &maps_in.m[ph], bar, p.in_c_off + kc * 8,
                          xy.tx * TW - p.org, xy.ty * TH - p.org, xy.img);
        }
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(kb.os_w[os]);
        for (int b = 0; b < kb.n_blocks; ++b, ++wc) {
          const int slot = wc % KB_WSLOTS;
          const int nm = min(kb.g_mma, p.n_mma - b * kb.g_mma);
          const uint32_t bytes = (uint32_t)nm * 2u * (uint32_t)p.nb * 16u;
          mbar_wait(wempty_bar0 + 8 * slot, ((wc / KB_WSLOTS) & 1u) ^ 1u);
          mbar_arrive_expect_tx(wfull_bar0 + 8 * slot, bytes);
          bulk_g2s(smem_u32(w_s) + (uint32_t)slot * kb.wslot_bytes, wsrc + (size_t)b * kb.wslot_bytes, bytes, wfull_bar0 + 8 * slot);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer: block-outer / tile-inner ===========================
    const uint32_t idesc = make_idesc_f16((uint32_t)p.nb);
    const uint32_t tbase = bcast0(tmem_base);
    const uint32_t a_hi = (((uint32_t)p.halo_w * 16u) >> 4) | (1u << 14);
    const uint32_t b_hi = (128u >> 4) | (1u << 14);
    const uint32_t b_step = 2u * (uint32_t)p.nb;
    uint32_t it = 0, wc = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int group = item / kb.n_os;
      const int nt = min(kb.T, p.n_tiles - group * kb.T);
      for (int b = 0; b < kb.n_blocks; ++b, ++wc) {
        const int slot = wc % KB_WSLOTS;
        const int j0 = b * kb.g_mma, nm = min(kb.g_mma, p.n_mma - j0);
        mbar_wait(wfull_bar0 + 8 * slot, (wc / KB_WSLOTS) & 1u);
        const uint32_t b_lo0 = ((smem_u32(w_s) + (uint32_t)slot * kb.wslot_bytes) >> 4) | ((((uint32_t)p.nb * 16u) >> 4) << 16);
        for (int t = 0; t < nt; ++t) {
          if (b == 0) {
            mbar_wait(full_bar0 + 8 * t, it & 1u);
            mbar_wait(acce_bar0 + 8 * t, (it & 1u) ^ 1u);
          }
          tc_fence_after();
          const uint32_t tacc = tbase + (uint32_t)(t * p.nb);
          const uint32_t a_base = (stage0 + (uint32_t)t * stage_bytes) >> 4;
          if (elect_one()) {
            for (int j = 0; j < nm; ++j) {
              const uint64_t adesc = ((uint64_t)a_hi << 32) | (uint64_t)(p.a_lo[j0 + j] + a_base);
              const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo0 + (uint32_t)j * b_step);
              umma_bf16_ss(tacc, adesc, bdesc, idesc, (b > 0 || j > 0) ? 1u : 0u);
            }
          }
          __syncwarp();
        }
        if (elect_one()) {
          umma_commit(wempty_bar0 + 8 * slot);
          if (b == kb.n_blocks - 1)
            for (int t = 0; t < nt; ++t) {
              umma_commit(empty_bar0 + 8 * t);
              umma_commit(accf_bar0 + 8 * t);
            }
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue: team (t & 1) takes tile t of every item ===========================
    const int ew = warp - EPI_WARP0;
    const int q = warp & 3;
    const int team = ew >> 2;
    const int m = q * 32 + lane;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int group = item / kb.n_os, os = item - group * kb.n_os;
      const int flags = kb.os_flags[os], out_c_off = kb.os_out_off[os];
      const bool has_res = (flags & PASS_RES_EXT) != 0, relu = (flags & PASS_RELU) != 0;
      const int nt = min(kb.T, p.n_tiles - group * kb.T);
      for (int t = team; t < nt; t += 2) {
        const int tile = group * kb.T + t;
        const uint32_t trow = tmem_base + (uint32_t)(t * p.nb) + ((uint32_t)(q * 32) << 16);
        const TileXY xy = tile_coords(tile, p);
        const int x = xy.tx * TW + (m & (TW - 1)), y = xy.ty * TH + (m >> 3);
        const bool live = x < p.W && y < p.H;
        const size_t pix = ((size_t)xy.img * p.H + (live ? y : 0)) * p.W + (live ? x : 0);
        DirectRes rr;
        if (has_res) direct_prefetch(p, rr, kb.res, pix, kb.out_ct, out_c_off, live);      // in flight while the MMAs finish
        mbar_wait(accf_bar0 + 8 * t, it & 1u);
        tc_fence_after();
        epilogue_rows_direct(p, rr, has_res, kb.out32, kb.out16, pix, kb.out_ct, out_c_off, live, trow, lo_col, relu,
                             bias_s + os * p.n_pad);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acce_bar0 + 8 * t);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, tmem_cols);
}

// BatchNorm-folded fp32 weights [cout_p][cin_p][k][k] (k*k = ntap) -> fp16 [n_mma][2][n_pad][8] in the kernel's K-list order
// split: [n_mma][2][2 n_pad][8], rows [0, n_pad) = fp16(w), rows [n_pad, 2 n_pad) = fp16(w - fp16(w)) (w_hi + w_lo = w to 2^-22)
__global__ void pack_kernel(const float* __restrict__ w, int cout, int cin, int kcg, int ntap, int nphase, int n_mma, int n_pad,
                            int split, __half* __restrict__ out) {
  const int nb = split ? 2 * n_pad : n_pad;
  const int total = n_mma * 2 * nb * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7, row = (i >> 3) % nb, slot = ((i >> 3) / nb) & 1, j = (i >> 3) / nb / 2;
    const int n = row < n_pad ? row : row - n_pad;
    int e0, e1;
    mma_entries(j, kcg, ntap, nphase, e0, e1);
    const int c = slot ? e1 : e0;
    float v = 0.f;
    if (c < ntap * kcg && n < cout) {
      const int tap = c / kcg, ch = (c - tap * kcg) * 8 + e;
      if (ch < cin) v = w[((size_t)n * cin + ch) * ntap + tap];
    }
    if (row >= n_pad) v = v - __half2float(__float2half_rn(v));
    out[i] = __float2half_rn(v);
  }
}

struct Geom {
  int kcg, kc, n_mma, n_pad, nphase, kdata, ch_stride;
};
static Geom geom(int cin, int cout, int ntap, int stride = 1) {
  Geom g;
  g.kcg = cin / 8;
  g.nphase = stride == 2 ? 4 : 1;
  g.kdata = g.nphase * g.kcg;
  g.kc = g.kdata + ((ntap * g.kcg) & 1);          // an odd K list ends on the all-zero chunk
  g.n_mma = (ntap * g.kcg + 1) / 2;
  g.n_pad = (cout + 15) / 16 * 16;
  g.ch_stride = ch_stride_of(ntap);
  return g;
}
static size_t smem_for(const Geom& g, int cout, int nbuf, int dual = 0, int split = 0) {
  return (size_t)nbuf * ((size_t)g.kc * g.ch_stride + 128 * (size_t)cout * (dual ? 6 : 2)) +
         (size_t)g.n_mma * 2 * g.n_pad * (split ? 2 : 1) * 16 + (size_t)g.n_pad * 4 + (2 * MAX_BUF + 8) * 8 + 16 + 128;
}

// ---- tensor maps (driver entry point fetched through the runtime: no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// [n_img, H, W, C] channels-last viewed as the 4-D tensor (C, W, H, n_img) with a (bc, bw, bh, 1) box.  phase >= 0: the
// (row parity py = phase >> 1, column parity px = phase & 1) plane of the map - the same memory with doubled pixel strides.
static int make_map(CUtensorMap* m, const void* base, int C, int W, int H, int n_img, int bc, int bw, int bh, int es_bytes = 2,
                    int phase = -1) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return LS3D_ERR_ARG;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img};
  cuuint64_t strides[3] = {(cuuint64_t)C * es_bytes, (cuuint64_t)W * C * es_bytes, (cuuint64_t)H * W * C * es_bytes};
  const char* b = (const char*)base;
  if (phase >= 0) {
    const int py = phase >> 1, px = phase & 1;
    dims[1] = (cuuint64_t)((W - px + 1) / 2);
    dims[2] = (cuuint64_t)((H - py + 1) / 2);
    if (dims[1] == 0 || dims[2] == 0) return LS3D_ERR_ARG;      // W or H == 1: no odd plane (rejected by the caller)
    strides[0] *= 2;
    strides[1] *= 2;
    b += ((size_t)py * W + px) * C * es_bytes;
  }
  const cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult r = fn(m, es_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<char*>(b), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LS3D_OK : 1000 + (int)r;
}

}  // namespace c3
}  // namespace ls3d

static int ntap_of(int ksize) { return ksize == 3 ? 9 : ksize == 1 ? 1 : 0; }
static bool shape_ok(int cin, int cout, int ksize, int stride) {
  return cin > 0 && cout > 0 && !(cin & 7) && !(cout & 7) && ntap_of(ksize) && (stride == 1 || (stride == 2 && ksize == 3));
}

extern "C" int ls3d_conv_f16_smem_bytes(int32_t cin, int32_t cout, int32_t ksize, int64_t* bytes) {
  using namespace ls3d::c3;
  if (!bytes || !shape_ok(cin, cout, ksize, 1)) return LS3D_ERR_ARG;
  *bytes = (int64_t)smem_for(geom(cin, cout, ntap_of(ksize)), cout, 2);
  return LS3D_OK;
}

extern "C" int ls3d_conv_f16_packed_bytes(int32_t cin, int32_t cout, int32_t ksize, int64_t* bytes) {
  using namespace ls3d::c3;
  if (!bytes || !shape_ok(cin, cout, ksize, 1)) return LS3D_ERR_ARG;
  const Geom g = geom(cin, cout, ntap_of(ksize));
  *bytes = (int64_t)g.n_mma * 2 * g.n_pad * 16;
  return LS3D_OK;
}

static int pack_launch(const float* w_oihw, int32_t cin, int32_t cout, int32_t ksize, int stride, int split, void* packed,
                       void* stream) {
  using namespace ls3d::c3;
  if (!w_oihw || !packed || !shape_ok(cin, cout, ksize, stride)) return LS3D_ERR_ARG;
  const Geom g = geom(cin, cout, ntap_of(ksize), stride);
  if (split && 2 * g.n_pad > 256) return LS3D_ERR_ARG;
  const int total = g.n_mma * 2 * g.n_pad * (split ? 2 : 1) * 8;
  pack_kernel<<<ls3d_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, cout, cin, g.kcg, ntap_of(ksize), g.nphase,
                                                                        g.n_mma, g.n_pad, split, (__half*)packed);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

extern "C" int ls3d_conv_f16_pack(const float* w_oihw, int32_t cin, int32_t cout, int32_t ksize, void* packed, void* stream) {
  return pack_launch(w_oihw, cin, cout, ksize, 1, 0, packed, stream);
}
extern "C" int ls3d_conv_f16_pack_split(const float* w_oihw, int32_t cin, int32_t cout, int32_t ksize, void* packed, void* stream) {
  return pack_launch(w_oihw, cin, cout, ksize, 1, 1, packed, stream);
}
// any (stride, split); packed holds ls3d_conv_f16_packed_bytes x (split ? 2 : 1)
extern "C" int ls3d_conv_f16_pack_ex(const float* w_oihw, int32_t cin, int32_t cout, int32_t ksize, int32_t stride, int32_t split,
                                     void* packed, void* stream) {
  return pack_launch(w_oihw, cin, cout, ksize, stride, split ? 1 : 0, packed, stream);
}

// does the kernel have a shared-memory / tensor-memory configuration for this launch shape?
extern "C" int ls3d_conv_f16_ex_supported(int32_t cin, int32_t cout, int32_t ksize, int32_t stride, int32_t dual, int32_t split,
                                          int32_t* supported) {
  using namespace ls3d::c3;
  if (!supported || !shape_ok(cin, cout, ksize, stride)) return LS3D_ERR_ARG;
  const Geom g = geom(cin, cout, ntap_of(ksize), stride);
  const int nb = split ? 2 * g.n_pad : g.n_pad;
  *supported = (nb <= 256 && g.n_mma <= MAX_MMA && smem_for(g, cout, 2, dual, split) <= 227 * 1024) ? 1 : 0;
  return LS3D_OK;
}
extern "C" int ls3d_conv_f16_split_supported(int32_t cin, int32_t cout, int32_t ksize, int32_t dual, int32_t* supported) {
  return ls3d_conv_f16_ex_supported(cin, cout, ksize, 1, dual, 1, supported);
}

static int conv_ex_launch(const ls3d_conv_args* c, const ls3d_conv_pass* passes, int n_pass, void* stream) {
  using namespace ls3d;
  using namespace ls3d::c3;
  if (!c || n_pass < 1 || n_pass > MAX_PASS || (n_pass > 1 && !passes)) return LS3D_ERR_ARG;
  if (c->n_img <= 0 || c->H_in <= 0 || c->W_in <= 0) return LS3D_OK;
  const int stride = c->stride ? c->stride : 1;
  if (!c->in16 || (!passes && !c->w_packed) || !shape_ok(c->cin, c->cout, c->ksize, stride)) return LS3D_ERR_ARG;
  const int dual = c->out32 != nullptr;
  void* out = dual ? (void*)c->out32 : c->out16;          // operand-only launches write the fp16 map alone
  if (!out || (dual && !c->out16) || (!dual && c->res32) || (dual && c->res16)) return LS3D_ERR_ARG;
  const void* res = dual ? (const void*)c->res32 : c->res16;
  const int in_ct = c->in_c_total ? c->in_c_total : c->cin, out_ct = c->out_c_total ? c->out_c_total : c->cout;
  if ((in_ct & 7) || (out_ct & 7)) return LS3D_ERR_ARG;
  for (int i = 0; i < n_pass; ++i) {
    const int io = passes ? passes[i].in_c_off : c->in_c_off, oo = passes ? passes[i].out_c_off : c->out_c_off;
    if ((io & 7) || (oo & 7) || io < 0 || oo < 0 || io + c->cin > in_ct || oo + c->cout > out_ct) return LS3D_ERR_ARG;
    if (passes && (!passes[i].w_packed || (((uintptr_t)passes[i].w_packed) & 15))) return LS3D_ERR_ARG;
    if (passes && (passes[i].flags & 4) && !dual) return LS3D_ERR_ARG;         // accumulation needs the fp32 output map
  }
  if ((((uintptr_t)c->in16) | ((uintptr_t)c->out32) | ((uintptr_t)res) | ((uintptr_t)c->w_packed) | ((uintptr_t)c->out16)) & 15)
    return LS3D_ERR_ARG;
  const int ntap = ntap_of(c->ksize), split = c->w_split ? 1 : 0;
  const int H = stride == 2 ? (c->H_in + 1) / 2 : c->H_in, W = stride == 2 ? (c->W_in + 1) / 2 : c->W_in;   // output size
  if (stride == 2 && (c->H_in < 2 || c->W_in < 2)) return LS3D_ERR_ARG;
  const Geom g = geom(c->cin, c->cout, ntap, stride);
  Args a;
  a.w = (const __half*)(passes ? passes[0].w_packed : c->w_packed); a.bias = c->bias; a.has_res = res != nullptr; a.dual = dual;
  a.n_img = c->n_img; a.H = H; a.W = W; a.cin = c->cin; a.cout = c->cout; a.relu = c->relu;
  a.n_pass = passes ? n_pass : 1;
  a.use_table = passes ? 1 : 0;
  if (passes) {
    // a single-entry pass list still goes through the table (flags), so force the multi-pass decode with n_pass >= 2 semantics
    for (int i = 0; i < n_pass; ++i) {
      a.pass_w[i] = (const __half*)passes[i].w_packed;
      a.pass_in_off[i] = (short)passes[i].in_c_off;
      a.pass_out_off[i] = (short)passes[i].out_c_off;
      int f = passes[i].flags & 15;
      if ((f & 2) && !res) f &= ~2;
      a.pass_flags[i] = (unsigned char)f;
    }
  }
  a.kcg = g.kcg; a.kc = g.kc; a.n_mma = g.n_mma; a.n_pad = g.n_pad; a.nphase = g.nphase; a.kdata = g.kdata;
  a.in_c_off = c->in_c_off; a.out_c_off = c->out_c_off;
  const int bw = ntap == 1 ? TW : HALO_W, bh = ntap == 1 ? TH : HALO_H;
  a.halo_w = bw; a.halo_pix = bw * bh; a.org = ntap == 1 ? 0 : 1; a.ch_stride = g.ch_stride;
  a.nb = split ? 2 * g.n_pad : g.n_pad;
  if (a.nb > 256 || g.n_mma > MAX_MMA) return LS3D_ERR_ARG;
  for (int j = 0; j < g.n_mma; ++j) {
    int e0, e1;
    mma_entries(j, g.kcg, ntap, g.nphase, e0, e1);
    const int o0 = k_entry_offset(e0, g.kcg, ntap, g.nphase), o1 = k_entry_offset(e1, g.kcg, ntap, g.nphase);
    a.a_lo[j] = ((uint32_t)o0 >> 4) | (((uint32_t)(o1 - o0) >> 4) << 16);
  }
  a.tiles_x = ls3d_div_up(W, TW);
  a.tiles_y = ls3d_div_up(H, TH);
  const long long nt = (long long)c->n_img * a.tiles_x * a.tiles_y;
  if (nt >= (1LL << 22)) return LS3D_ERR_ARG;
  a.n_tiles = (int)nt;
  a.inv_tiles_x = 1.0f / (float)a.tiles_x;
  a.inv_tiles_per_img = 1.0f / (float)(a.tiles_x * a.tiles_y);
  a.nbuf = MAX_BUF;
  while (a.nbuf > 2 && smem_for(g, c->cout, a.nbuf, dual, split) > 227 * 1024) --a.nbuf;
  const size_t smem = smem_for(g, c->cout, a.nbuf, dual, split);
  if (smem > 227 * 1024) return LS3D_ERR_ARG;           // weights do not fit in shared memory: slice the channels / library conv
  const int num_sms = ls3d_num_sms();
  static bool optin[64] = {false};
  cudaError_t eo = ls3d_optin_smem(conv3x3_f16_kernel, optin);
  if (eo != cudaSuccess) return (int)eo;
  InMaps m_in;
  CUtensorMap m_res, m_out, m_out16;
  const int es = dual ? 4 : 2;
  int rc;
  for (int ph = 0; ph < 4; ++ph) {
    rc = make_map(&m_in.m[ph], c->in16, in_ct, c->W_in, c->H_in, c->n_img, 8, bw, bh, 2, stride == 2 ? ph : -1);
    if (rc) return rc;
    if (stride != 2 && ph == 0) {
      m_in.m[1] = m_in.m[2] = m_in.m[3] = m_in.m[0];
      break;
    }
  }
  rc = make_map(&m_out, out, out_ct, W, H, c->n_img, c->cout, TW, TH, es);
  if (rc) return rc;
  rc = make_map(&m_res, res ? res : out, out_ct, W, H, c->n_img, c->cout, TW, TH, es);
  if (rc) return rc;
  rc = make_map(&m_out16, dual ? c->out16 : out, out_ct, W, H, c->n_img, c->cout, TW, TH, 2);
  if (rc) return rc;
  const int grid = a.n_tiles < num_sms ? a.n_tiles : num_sms;
  conv3x3_f16_kernel<<<grid, N_THREADS, smem, (cudaStream_t)stream>>>(a, m_in, m_res, m_out, m_out16);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

// ---- K-block (streamed-weight) variant: configuration for a launch shape, or T = 0 when it does not apply
struct KbCfg {
  int T, g_mma, n_blocks;
  size_t smem;
};
static KbCfg kb_config(int cin, int cout, int dual, int split, long long n_tiles, int num_sms, int stride = 1, int ntap = 9) {
  using namespace ls3d::c3;
  KbCfg k = {0, 0, 0, 0};
  const Geom g = geom(cin, cout, ntap, stride);
  const int nb = split ? 2 * g.n_pad : g.n_pad;
  if (nb > 256 || g.n_mma > MAX_MMA) return k;
  (void)dual;
  if (cout > DIRECT_MAX_C) return k;
  const size_t stage = (size_t)g.kc * g.ch_stride;                 // halo only: the K-block kernel's epilogue goes straight to global memory
  const size_t fixed = (size_t)8 * g.n_pad * 4 + (4 * MAX_BUF + 2 * KB_WSLOTS) * 8 + 16 + 128;
  const size_t per_mma = (size_t)2 * nb * 16;
  int T = (int)((n_tiles + num_sms - 1) / num_sms);
  if (T > 512 / nb) T = 512 / nb;
  if (T > 4) T = 4;
  for (; T >= 1; --T) {
    const size_t left = 227 * 1024 - fixed;
    if ((size_t)T * stage + KB_WSLOTS * 2 * per_mma > left) continue;
    int gm = (int)((left - (size_t)T * stage) / (KB_WSLOTS * per_mma));
    if (gm > g.n_mma) gm = g.n_mma;
    if (gm < 2) continue;
    k.T = T;
    k.g_mma = gm;
    k.n_blocks = (g.n_mma + gm - 1) / gm;
    k.smem = (size_t)T * stage + KB_WSLOTS * (size_t)gm * per_mma + fixed;
    return k;
  }
  return k;
}

extern "C" int ls3d_conv_f16_kb_supported(int32_t cin, int32_t cout, int32_t ksize, int32_t stride, int32_t dual, int32_t split,
                                          int64_t n_out_pixels, int32_t* supported) {
  if (!supported || (stride != 1 && stride != 2) || !shape_ok(cin, cout, ksize, stride)) return LS3D_ERR_ARG;
  const KbCfg k = kb_config(cin, cout, dual, split, (n_out_pixels + 127) / 128, ls3d_num_sms() > 0 ? ls3d_num_sms() : 148, stride,
                            ntap_of(ksize));
  *supported = k.T > 0 ? 1 : 0;
  return LS3D_OK;
}

// 3x3 / stride 1 convolution with STREAMED weights (conv3x3_kb_kernel): args as ls3d_conv_f16_multi, every pass = one output
// channel slice over ALL input channels (pass.in_c_off must be 0, no accumulating passes); args->cout = slice size.
extern "C" int ls3d_conv_f16_kb(const ls3d_conv_args* c, const ls3d_conv_pass* passes, int32_t n_os, void* stream) {
  using namespace ls3d;
  using namespace ls3d::c3;
  if (!c || !passes || n_os < 1 || n_os > 8) return LS3D_ERR_ARG;
  if (c->n_img <= 0 || c->H_in <= 0 || c->W_in <= 0) return LS3D_OK;
  const int stride = c->stride ? c->stride : 1;
  if (!c->in16 || (stride != 1 && stride != 2) || !shape_ok(c->cin, c->cout, c->ksize, stride)) return LS3D_ERR_ARG;
  const int ntap = ntap_of(c->ksize);
  if (stride == 2 && (c->H_in < 2 || c->W_in < 2)) return LS3D_ERR_ARG;
  const int dual = c->out32 != nullptr;
  void* out = dual ? (void*)c->out32 : c->out16;
  if (!out || (dual && !c->out16) || (!dual && c->res32) || (dual && c->res16)) return LS3D_ERR_ARG;
  const void* res = dual ? (const void*)c->res32 : c->res16;
  const int in_ct = c->in_c_total ? c->in_c_total : c->cin, out_ct = c->out_c_total ? c->out_c_total : c->cout;
  if ((in_ct & 7) || (out_ct & 7) || in_ct != c->cin) return LS3D_ERR_ARG;
  if ((((uintptr_t)c->in16) | ((uintptr_t)c->out32) | ((uintptr_t)res) | ((uintptr_t)c->out16)) & 15) return LS3D_ERR_ARG;
  const int split = c->w_split ? 1 : 0;
  const int H = stride == 2 ? (c->H_in + 1) / 2 : c->H_in, W = stride == 2 ? (c->W_in + 1) / 2 : c->W_in;   // output size
  const Geom g = geom(c->cin, c->cout, ntap, stride);
  Args a = {};
  KbArgs kb = {};
  for (int i = 0; i < n_os; ++i) {
    if (!passes[i].w_packed || (((uintptr_t)passes[i].w_packed) & 15) || passes[i].in_c_off != 0 || (passes[i].out_c_off & 7) ||
        passes[i].out_c_off < 0 || passes[i].out_c_off + c->cout > out_ct || (passes[i].flags & 4))
      return LS3D_ERR_ARG;
    kb.os_w[i] = (const __half*)passes[i].w_packed;
    kb.os_out_off[i] = (short)passes[i].out_c_off;
    int f = passes[i].flags & 15;
    if ((f & 2) && !res) f &= ~2;
    kb.os_flags[i] = (unsigned char)f;
  }
  a.bias = c->bias; a.dual = dual; a.n_img = c->n_img; a.H = H; a.W = W; a.cin = c->cin; a.cout = c->cout;
  a.kcg = g.kcg; a.kc = g.kc; a.n_mma = g.n_mma; a.n_pad = g.n_pad; a.nphase = g.nphase; a.kdata = g.kdata;
  a.in_c_off = 0; a.out_c_off = 0; a.n_pass = 1; a.use_table = 0;
  const int bw = ntap == 1 ? TW : HALO_W, bh = ntap == 1 ? TH : HALO_H;
  a.halo_w = bw; a.halo_pix = bw * bh; a.org = ntap == 1 ? 0 : 1; a.ch_stride = g.ch_stride;
  a.nb = split ? 2 * g.n_pad : g.n_pad;
  if (a.nb > 256 || g.n_mma > MAX_MMA) return LS3D_ERR_ARG;
  for (int j = 0; j < g.n_mma; ++j) {
    int e0, e1;
    mma_entries(j, g.kcg, ntap, g.nphase, e0, e1);
    const int o0 = k_entry_offset(e0, g.kcg, ntap, g.nphase), o1 = k_entry_offset(e1, g.kcg, ntap, g.nphase);
    a.a_lo[j] = ((uint32_t)o0 >> 4) | (((uint32_t)(o1 - o0) >> 4) << 16);
  }
  a.tiles_x = ls3d_div_up(W, TW);
  a.tiles_y = ls3d_div_up(H, TH);
  const long long nt = (long long)c->n_img * a.tiles_x * a.tiles_y;
  if (nt >= (1LL << 22)) return LS3D_ERR_ARG;
  a.n_tiles = (int)nt;
  a.inv_tiles_x = 1.0f / (float)a.tiles_x;
  a.inv_tiles_per_img = 1.0f / (float)(a.tiles_x * a.tiles_y);
  const int num_sms = ls3d_num_sms();
  const KbCfg k = kb_config(c->cin, c->cout, dual, split, nt, num_sms, stride, ntap);
  if (k.T < 1) return LS3D_ERR_ARG;
  a.nbuf = k.T;
  kb.T = k.T; kb.g_mma = k.g_mma; kb.n_blocks = k.n_blocks; kb.n_os = n_os;
  kb.n_groups = (int)((nt + k.T - 1) / k.T);
  kb.wslot_bytes = (uint32_t)k.g_mma * 2u * (uint32_t)a.nb * 16u;
  kb.res = res; kb.out32 = dual ? c->out32 : nullptr; kb.out16 = (__half*)(dual ? c->out16 : out); kb.out_ct = out_ct;
  static bool optin[64] = {false};
  cudaError_t eo = ls3d_optin_smem(conv3x3_kb_kernel, optin);
  if (eo != cudaSuccess) return (int)eo;
  InMaps m_in;
  CUtensorMap m_res, m_out, m_out16;
  const int es = dual ? 4 : 2;
  int rc = 0;
  for (int ph = 0; ph < 4; ++ph) {
    rc = make_map(&m_in.m[ph], c->in16, in_ct, c->W_in, c->H_in, c->n_img, 8, bw, bh, 2, stride == 2 ? ph : -1);
    if (rc) return rc;
    if (stride != 2 && ph == 0) {
      m_in.m[1] = m_in.m[2] = m_in.m[3] = m_in.m[0];
      break;
    }
  }
  (void)es;
  m_res = m_out = m_out16 = m_in.m[0];          // the K-block kernel's epilogue addresses the output maps directly
  const int n_items = kb.n_groups * n_os;
  const int grid = n_items < num_sms ? n_items : num_sms;
  conv3x3_kb_kernel<<<grid, N_THREADS, k.smem, (cudaStream_t)stream>>>(a, kb, m_in, m_res, m_out, m_out16);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

extern "C" int ls3d_conv_f16_ex(const ls3d_conv_args* c, void* stream) { return conv_ex_launch(c, nullptr, 1, stream); }

// several passes (weight block, input / output channel slice, flags) over the same tiles in ONE launch; args->cin / cout are the
// uniform slice sizes, args->bias the bias of the whole output tensor, args->in_c_off / out_c_off / relu / w_packed are unused
extern "C" int ls3d_conv_f16_multi(const ls3d_conv_args* c, const ls3d_conv_pass* passes, int32_t n_pass, void* stream) {
  if (!passes) return LS3D_ERR_ARG;
  return conv_ex_launch(c, passes, n_pass, stream);
}

// fp16 maps (round-1 entry point): fp16 residual / output
extern "C" int ls3d_conv_f16(const void* in, const void* w_packed, const float* bias, const void* res, void* out, int32_t n_img,
                             int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t ksize, int32_t relu, void* stream) {
  using namespace ls3d;
  using namespace ls3d::c3;
  if (n_img <= 0 || H <= 0 || W <= 0) return LS3D_OK;
  if (!in || !w_packed || !out || !shape_ok(cin, cout, ksize, 1)) return LS3D_ERR_ARG;
  if ((((uintptr_t)in) | ((uintptr_t)out) | ((uintptr_t)res) | ((uintptr_t)w_packed)) & 15) return LS3D_ERR_ARG;
  const int ntap = ntap_of(ksize);
  const Geom g = geom(cin, cout, ntap);
  Args a;
  a.w = (const __half*)w_packed; a.bias = bias; a.has_res = res != nullptr; a.dual = 0;
  a.n_img = n_img; a.H = H; a.W = W; a.cin = cin; a.cout = cout; a.relu = relu;
  a.kcg = g.kcg; a.kc = g.kc; a.n_mma = g.n_mma; a.n_pad = g.n_pad; a.nphase = 1; a.kdata = g.kdata;
  a.in_c_off = a.out_c_off = 0;
  a.n_pass = 1;
  a.use_table = 0;
  const int bw = ntap == 1 ? TW : HALO_W, bh = ntap == 1 ? TH : HALO_H;
  a.halo_w = bw; a.halo_pix = bw * bh; a.org = ntap == 1 ? 0 : 1; a.ch_stride = g.ch_stride;
  a.nb = g.n_pad;
  if (a.nb > 256 || g.n_mma > MAX_MMA) return LS3D_ERR_ARG;
  for (int j = 0; j < g.n_mma; ++j) {
    int e0, e1;
    mma_entries(j, g.kcg, ntap, 1, e0, e1);
    const int o0 = k_entry_offset(e0, g.kcg, ntap), o1 = k_entry_offset(e1, g.kcg, ntap);
    a.a_lo[j] = ((uint32_t)o0 >> 4) | (((uint32_t)(o1 - o0) >> 4) << 16);
  }
  a.tiles_x = ls3d_div_up(W, TW);
  a.tiles_y = ls3d_div_up(H, TH);
  const long long nt = (long long)n_img * a.tiles_x * a.tiles_y;
  if (nt >= (1LL << 22)) return LS3D_ERR_ARG;
  a.n_tiles = (int)nt;
  a.inv_tiles_x = 1.0f / (float)a.tiles_x;
  a.inv_tiles_per_img = 1.0f / (float)(a.tiles_x * a.tiles_y);
  a.nbuf = MAX_BUF;
  while (a.nbuf > 2 && smem_for(g, cout, a.nbuf, 0, 0) > 227 * 1024) --a.nbuf;
  const size_t smem = smem_for(g, cout, a.nbuf, 0, 0);
  if (smem > 227 * 1024) return LS3D_ERR_ARG;
  const int num_sms = ls3d_num_sms();
  static bool optin[64] = {false};
  cudaError_t eo = ls3d_optin_smem(conv3x3_f16_kernel, optin);
  if (eo != cudaSuccess) return (int)eo;
  InMaps m_in;
  CUtensorMap m_res, m_out;
  int rc = make_map(&m_in.m[0], in, cin, W, H, n_img, 8, bw, bh);
  if (rc) return rc;
  m_in.m[1] = m_in.m[2] = m_in.m[3] = m_in.m[0];
  rc = make_map(&m_out, out, cout, W, H, n_img, cout, TW, TH, 2);
  if (rc) return rc;
  rc = make_map(&m_res, res ? res : out, cout, W, H, n_img, cout, TW, TH, 2);
  if (rc) return rc;
  const int grid = a.n_tiles < num_sms ? a.n_tiles : num_sms;
  conv3x3_f16_kernel<<<grid, N_THREADS, smem, (cudaStream_t)stream>>>(a, m_in, m_res, m_out, m_out);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

// fp32 feature maps with fp16 tensor-core operands (see include/ls3d.h); thin wrapper over ls3d_conv_f16_ex
extern "C" int ls3d_conv_f16_dual(const void* in16, const void* w_packed, const float* bias, const float* res32, float* out32,
                                  void* out16, int32_t n_img, int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t ksize,
                                  int32_t relu, int32_t w_split, void* stream) {
  ls3d_conv_args c = {};
  c.in16 = in16; c.in_c_total = cin; c.in_c_off = 0; c.cin = cin;
  c.w_packed = w_packed; c.bias = bias; c.res32 = res32; c.out32 = out32; c.out16 = out16;
  c.out_c_total = cout; c.out_c_off = 0; c.cout = cout;
  c.n_img = n_img; c.H_in = H; c.W_in = W; c.ksize = ksize; c.stride = 1; c.relu = relu; c.w_split = w_split;
  return ls3d_conv_f16_ex(&c, stream);
}

extern "C" int ls3d_conv_f16_dual_smem_bytes(int32_t cin, int32_t cout, int32_t ksize, int64_t* bytes) {
  using namespace ls3d::c3;
  if (!bytes || !shape_ok(cin, cout, ksize, 1)) return LS3D_ERR_ARG;
  *bytes = (int64_t)smem_for(geom(cin, cout, ntap_of(ksize)), cout, 2, 1);
  return LS3D_OK;
}

// the 3x3 entry points (ksize = 3)
extern "C" int ls3d_conv3x3_f16_smem_bytes(int32_t cin, int32_t cout, int64_t* bytes) {
  return ls3d_conv_f16_smem_bytes(cin, cout, 3, bytes);
}
extern "C" int ls3d_conv3x3_f16_packed_bytes(int32_t cin, int32_t cout, int64_t* bytes) {
  return ls3d_conv_f16_packed_bytes(cin, cout, 3, bytes);
}
extern "C" int ls3d_conv3x3_f16_pack(const float* w_oihw, int32_t cin, int32_t cout, void* packed, void* stream) {
  return ls3d_conv_f16_pack(w_oihw, cin, cout, 3, packed, stream);
}
extern "C" int ls3d_conv3x3_f16(const void* in, const void* w_packed, const float* bias, const void* res, void* out,
                                int32_t n_img, int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t relu, void* stream) {
  return ls3d_conv_f16(in, w_packed, bias, res, out, n_img, H, W, cin, cout, 3, relu, stream);
}
