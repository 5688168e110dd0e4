// Camera stem: fused 3x3 / stride 1 / pad 1 convolution + folded BatchNorm bias (+ residual) (+ ReLU) on channels-last fp16
// maps, tcgen05 tensor cores with fp32 accumulation in TMEM.
//
// Replaces (eval mode) the conv3x3 -> BN -> [+identity] -> ReLU pairs of the HRNet BasicBlocks, reference
// det3d/models/img_backbones/resnet_mmcv.py:20-100 as instantiated by hrnet.py:78-226 (4 blocks per branch and module; 64 of
// them per forward on the 18-channel full-resolution branch alone), which cuDNN serves with 55-130 us launches at 18-72
// channels.  fp16 operands carry the same 10-bit mantissa as the TF32 tensor-core path the fp32 stem uses.
//
// Implicit GEMM without im2col traffic: a persistent CTA walks 16 x 8 output tiles (128 pixels = the UMMA M dimension).
// The (16+2) x (8+2) input halo of a tile is staged ONCE in shared memory, channel-chunk major ([16-byte chunk][pixel]) -
// the un-swizzled K-major UMMA layout whose 8-row groups are 8 x-adjacent pixels - so each of the 9 taps is the same buffer
// read through a shared-memory descriptor whose start address is shifted by (dy * 10 + dx) pixels: 9 * Cin/16 MMAs per tile
// against weights that stay resident in shared memory for the whole launch.
//   warps 0-2   loaders : cp.async halo gathers (zero fill outside the image) driven by a per-launch shared-memory table of
//                         (global offset, halo position) per 16-byte chunk; completion via cp.async.mbarrier.arrive
//   warp 3      MMA     : one elected thread, tcgen05.mma.kind::f16 M=128 N=Cout K=16, double-buffered accumulators
//   warps 4-7   epilogue: tcgen05.ld -> + bias (+ residual) -> ReLU -> fp16 -> 16-byte stores; overlaps the next tile's MMAs
// Bound: HBM (one read of the input, one of the residual, one write) once the ~64-cycle per-MMA operand fetch is hidden.
#include <cuda_fp16.h>

#include "common.cuh"
#include "../../include/ls3d.h"

namespace ls3d {
namespace c3 {

constexpr int TW = 8, TH = 16;              // output tile: 8 wide x 16 tall
constexpr int HALO_W = TW + 2, HALO_H = TH + 2;
constexpr int HPIX = HALO_W * HALO_H;       // 180 halo pixels
constexpr int N_LOAD_WARPS = 3;              // warps 0-2 loaders, warp 3 MMA issuer, warps 4-7 epilogue
constexpr int N_LOAD = N_LOAD_WARPS * 32;
constexpr int MMA_WARP = N_LOAD_WARPS;
constexpr int N_THREADS = 8 * 32;
constexpr int MAX_BUF = 4;

struct Args {
  const __half* in;
  const __half* w;        // [9][k_pad/8][n_pad][8] fp16 (tap = ky*3+kx, 8 input channels per 16-byte chunk)
  const float* bias;      // [cout] fp32 or NULL
  const __half* res;      // [n_img,H,W,cout] or NULL
  __half* out;            // [n_img,H,W,cout]
  int n_img, H, W, cin, cout, k_pad, n_pad, relu, tiles_x, tiles_y, n_tiles, nbuf;
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

__global__ void __launch_bounds__(N_THREADS, 1) conv3x3_f16_kernel(const Args p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int KC = p.k_pad / 8;                             // 16-byte chunks along K (shared memory)
  const int KCG = p.cin / 8;                              // chunks present in global memory
  const uint32_t w_bytes = 9u * KC * p.n_pad * 16u;
  const uint32_t halo_bytes = (uint32_t)KC * HPIX * 16u;
  uint8_t* w_s = smem;
  uint8_t* halo_s = w_s + w_bytes;
  float* bias_s = (float*)(halo_s + p.nbuf * halo_bytes);
  int2* tab_s = (int2*)(bias_s + p.n_pad);                   // [HPIX * KCG] copy table (tile independent)
  uint64_t* bars = (uint64_t*)(tab_s + HPIX * KCG);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * MAX_BUF + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t full_bar0 = smem_u32(bars);                   // halo landed   [nbuf] (32 noinc arrivals)
  const uint32_t empty_bar0 = smem_u32(bars + MAX_BUF);        // halo consumed [nbuf] (tcgen05.commit)
  const uint32_t accf_bar0 = smem_u32(bars + 2 * MAX_BUF);     // accumulator full  [2]
  const uint32_t acce_bar0 = smem_u32(bars + 2 * MAX_BUF + 2); // accumulator empty [2]

  // ---- one-time staging: weights (resident), bias, zero K padding of the halo buffers
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(w_s);
    for (uint32_t i = tid; i < w_bytes / 16; i += N_THREADS) dst[i] = __ldg(src + i);
    for (int c = tid; c < p.n_pad; c += N_THREADS) bias_s[c] = (p.bias && c < p.cout) ? __ldg(p.bias + c) : 0.f;
    // copy i of a tile: 16-byte chunk kc of halo pixel (hy, hx); consecutive i = consecutive chunks of a pixel (coalesced).
    // .x = element offset from the tile's first halo pixel, .y = hy | hx << 8 | (chunk-major smem slot) << 16
    for (int i = tid; i < HPIX * KCG; i += N_THREADS) {
      const int pix = i / KCG, kc = i - pix * KCG;
      const int hy = pix / HALO_W, hx = pix - hy * HALO_W;
      tab_s[i] = make_int2((hy * p.W + hx) * p.cin + kc * 8, hy | (hx << 8) | ((kc * HPIX + pix) << 16));
    }
    if (KCG < KC) {
      uint4* h = reinterpret_cast<uint4*>(halo_s);
      const int per = (KC - KCG) * HPIX;
      for (int i = tid; i < p.nbuf * per; i += N_THREADS)
        h[(size_t)(i / per) * KC * HPIX + (size_t)KCG * HPIX + (i % per)] = make_uint4(0, 0, 0, 0);
    }
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * (uint32_t)p.n_pad) tmem_cols <<= 1;
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < MAX_BUF; ++s) {
        mbar_init(full_bar0 + 8 * s, N_LOAD);
        mbar_init(empty_bar0 + 8 * s, 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(accf_bar0 + 8 * b, 1);
        mbar_init(acce_bar0 + 8 * b, 128);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  }
  fence_proxy_async_smem();          // the staged weights / zero padding are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp < N_LOAD_WARPS) {
    // =========================== halo loaders ===========================
    int it = 0;
    const int n_copy = HPIX * KCG;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int b = it % p.nbuf;
      const uint32_t ph = (uint32_t)(it / p.nbuf) & 1u;
      mbar_wait(empty_bar0 + 8 * b, ph ^ 1u);
      const int img = tile / tiles_per_img, rem = tile % tiles_per_img;
      const int y0 = (rem / p.tiles_x) * TH - 1, x0 = (rem % p.tiles_x) * TW - 1;
      const uint32_t dst0 = smem_u32(halo_s + (size_t)b * halo_bytes);
      // first halo pixel of the tile (may lie outside the image: only dereferenced for in-image pixels)
      const __half* org = p.in + ((size_t)img * p.H * p.W + (long long)y0 * p.W + x0) * p.cin;
      for (int i = tid; i < n_copy; i += N_LOAD) {
        const int2 t = tab_s[i];
        const int gy = y0 + (t.y & 0xff), gx = x0 + ((t.y >> 8) & 0xff);
        const bool ok = (unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W;
        cp_async16(dst0 + ((uint32_t)t.y >> 16) * 16u, ok ? org + t.x : p.in, ok ? 16u : 0u);
      }
      cp_async_mbar_arrive_noinc(full_bar0 + 8 * b);
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc_f16((uint32_t)p.n_pad);
    const uint32_t tbase = bcast0(tmem_base);
    const uint32_t w0 = smem_u32(w_s), h0 = smem_u32(halo_s);
    const int nsl = p.k_pad / 16;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int b = it % p.nbuf;
      const int ab = it & 1;
      mbar_wait(full_bar0 + 8 * b, (uint32_t)(it / p.nbuf) & 1u);
      mbar_wait(acce_bar0 + 8 * ab, ((uint32_t)(it >> 1) & 1u) ^ 1u);
      fence_proxy_async_smem();                 // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
      tc_fence_after();
      const uint32_t tacc = tbase + (uint32_t)(ab * p.n_pad);
      const uint32_t hb = h0 + (uint32_t)b * halo_bytes;
      if (elect_one()) {
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3, dx = tap - dy * 3;
          for (int j = 0; j < nsl; ++j) {
            const uint64_t adesc = make_desc_k_nosw(hb + (uint32_t)((2 * j) * HPIX + dy * HALO_W + dx) * 16u, HPIX * 16u,
                                                    HALO_W * 16u);
            const uint64_t bdesc = make_desc_k_nosw(w0 + (uint32_t)((tap * KC + 2 * j) * p.n_pad) * 16u,
                                                    (uint32_t)p.n_pad * 16u, 128u);
            umma_bf16_ss(tacc, adesc, bdesc, idesc, (tap > 0 || j > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty_bar0 + 8 * b);
        umma_commit(accf_bar0 + 8 * ab);
      }
      __syncwarp();
    }
  } else {
    // =========================== epilogue ===========================
    const int q = warp & 3;                       // TMEM lane quarter of this warp
    const int m = q * 32 + lane;                  // tile pixel = accumulator row
    const int py = m >> 3, px = m & 7;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const int img = tile / tiles_per_img, rem = tile % tiles_per_img;
      const int gy = (rem / p.tiles_x) * TH + py, gx = (rem % p.tiles_x) * TW + px;
      const bool ok = gy < p.H && gx < p.W;
      const size_t pix = ((size_t)img * p.H + gy) * p.W + gx;
      mbar_wait(accf_bar0 + 8 * ab, (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const uint32_t trow = tmem_base + (uint32_t)(ab * p.n_pad) + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
        uint32_t raw[16];
        tmem_ld16(trow + c0, raw);
        uint4 rz[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        if (p.res && ok) {
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8)
            if (c0 + 8 * g8 < p.cout) rz[g8] = __ldg(reinterpret_cast<const uint4*>(p.res + pix * p.cout + c0 + 8 * g8));
        }
        tmem_ld_wait();
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          if (c0 + 8 * g8 >= p.cout) continue;
          const __half2* r2 = reinterpret_cast<const __half2*>(&rz[g8]);
          uint4 o;
          __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c0 + 8 * g8 + 2 * e;
            const float2 rr = __half22float2(r2[e]);
            float v0 = __uint_as_float(raw[8 * g8 + 2 * e]) + bias_s[c] + rr.x;
            float v1 = __uint_as_float(raw[8 * g8 + 2 * e + 1]) + bias_s[c + 1] + rr.y;
            if (p.relu) {
              v0 = fmaxf(v0, 0.f);
              v1 = fmaxf(v1, 0.f);
            }
            o2[e] = __floats2half2_rn(v0, v1);
          }
          if (ok) *reinterpret_cast<uint4*>(p.out + pix * p.cout + c0 + 8 * g8) = o;
        }
      }
      tc_fence_before();
      mbar_arrive(acce_bar0 + 8 * ab);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, tmem_cols);
}

static size_t smem_for(int k_pad, int n_pad, int nbuf) {
  const size_t KC = k_pad / 8;
  return 9 * KC * n_pad * 16 + (size_t)nbuf * KC * HPIX * 16 + (size_t)n_pad * 4 + (size_t)HPIX * KC * 8 +
         (2 * MAX_BUF + 4) * 8 + 16 + 128;
}

}  // namespace c3
}  // namespace ls3d

extern "C" int ls3d_conv3x3_f16_smem_bytes(int32_t cin, int32_t cout, int64_t* bytes) {
  using namespace ls3d::c3;
  if (!bytes || cin <= 0 || cout <= 0) return LS3D_ERR_ARG;
  *bytes = (int64_t)smem_for((cin + 15) / 16 * 16, (cout + 15) / 16 * 16, 2);
  return LS3D_OK;
}

extern "C" int ls3d_conv3x3_f16(const void* in, const void* w_packed, const float* bias, const void* res, void* out,
                                int32_t n_img, int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t relu, void* stream) {
  using namespace ls3d;
  using namespace ls3d::c3;
  if (n_img <= 0 || H <= 0 || W <= 0) return LS3D_OK;
  if (!in || !w_packed || !out || cin <= 0 || cout <= 0 || (cin & 7) || (cout & 7)) return LS3D_ERR_ARG;
  Args a;
  a.in = (const __half*)in; a.w = (const __half*)w_packed; a.bias = bias; a.res = (const __half*)res; a.out = (__half*)out;
  a.n_img = n_img; a.H = H; a.W = W; a.cin = cin; a.cout = cout; a.relu = relu;
  a.k_pad = (cin + 15) / 16 * 16;
  a.n_pad = (cout + 15) / 16 * 16;
  if (a.n_pad > 256) return LS3D_ERR_ARG;
  a.tiles_x = ls3d_div_up(W, TW);
  a.tiles_y = ls3d_div_up(H, TH);
  const long long nt = (long long)n_img * a.tiles_x * a.tiles_y;
  if (nt > 0x7fffffffLL) return LS3D_ERR_ARG;
  a.n_tiles = (int)nt;
  a.nbuf = MAX_BUF;
  while (a.nbuf > 2 && smem_for(a.k_pad, a.n_pad, a.nbuf) > 227 * 1024) --a.nbuf;
  const size_t smem = smem_for(a.k_pad, a.n_pad, a.nbuf);
  if (smem > 227 * 1024) return LS3D_ERR_ARG;           // weights do not fit in shared memory: caller uses the library conv
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
    cudaError_t e = cudaFuncSetAttribute(conv3x3_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
  }
  const int grid = a.n_tiles < num_sms ? a.n_tiles : num_sms;
  conv3x3_f16_kernel<<<grid, N_THREADS, smem, (cudaStream_t)stream>>>(a);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}
