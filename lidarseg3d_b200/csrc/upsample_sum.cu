// Multi-resolution branch fusion of the camera stem: out = act(bias + sum_k resize(term_k)) in ONE pass over channels-last maps
// (bias = the summed folded-BatchNorm shifts of the bias-free convolutions that produced the terms: a per-channel constant
// commutes with the bilinear resize, whose taps sum to one).
//
// Replaces the python loops of HRModule.forward (reference det3d/models/img_backbones/hrnet.py:205-226: per output branch,
// y += x_i | y += resize(conv1x1_bn(x_j)) for coarser j | y += strided convs of finer j, then ReLU) and the
// resize_concat + first 1x1 ConvModule of the FCN decode head (det3d/models/img_heads/decode_head.py:151-160,
// fcn_mseg3d_head.py:150-163; the 1x1 convolution commutes with the bilinear resize, so the per-branch convolutions run at
// native resolution and only their sum is formed at full resolution).  The reference issues one bilinear-upsample kernel,
// one add and one ReLU per term; here every output element is written once and the coarse terms (<= 1/4 of the pixels)
// are read through L1/L2.  Bilinear taps follow ATen's align_corners=False rule (src = scale*(dst+0.5)-0.5, clamped at 0).
// One thread per (pixel, 4 channels): 16-byte loads/stores; HBM-bound: (1 + sum_k hk*wk/(H*W)) reads + 1 write per element.
#include <cuda_fp16.h>

#include "common.cuh"
#include "../../include/ls3d.h"

namespace ls3d {

constexpr int UPS_MAX_TERMS = 4;

struct UpsTerms {
  const float* p[UPS_MAX_TERMS];
  int h[UPS_MAX_TERMS], w[UPS_MAX_TERMS];
  float rh[UPS_MAX_TERMS], rw[UPS_MAX_TERMS];
  int n;
};

__device__ __forceinline__ float4 f4_fma(float a, float4 x, float4 acc) {
  return make_float4(fmaf(a, x.x, acc.x), fmaf(a, x.y, acc.y), fmaf(a, x.z, acc.z), fmaf(a, x.w, acc.w));
}
__device__ __forceinline__ float4 f4_scale(float a, float4 x) { return make_float4(a * x.x, a * x.y, a * x.z, a * x.w); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__global__ void __launch_bounds__(256) upsample_sum_kernel(UpsTerms T, int n_img, int H, int W, int C4, int relu,
                                                           const float* __restrict__ bias, float* __restrict__ out,
                                                           uint2* __restrict__ out16) {
  // grid = (x / channel-group blocks, rows, images): the row and the image come from the block index and the only division
  // left per thread is a 32-bit one by C4 (the flat 64-bit index arithmetic of the first version cost more issue slots than
  // the whole resize: 65-77 us per full-resolution fusion against a 25 us HBM floor)
  const int y = blockIdx.y, img = blockIdx.z;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < W * C4; t += gridDim.x * blockDim.x) {
    const int x = t / C4, c4 = t - x * C4;
    const size_t e = ((size_t)img * H + y) * (size_t)W * C4 + t;
    float4 acc = bias ? __ldg(reinterpret_cast<const float4*>(bias) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < UPS_MAX_TERMS; ++k) {
      if (k >= T.n) break;
      const int h = T.h[k], w = T.w[k];
      const float4* src = reinterpret_cast<const float4*>(T.p[k]) + (size_t)img * h * w * C4 + c4;
      float4 v;
      if (h == H && w == W) {
        v = __ldg(src + ((size_t)y * W + x) * C4);
      } else {
        float sy = T.rh[k] * ((float)y + 0.5f) - 0.5f;
        float sx = T.rw[k] * ((float)x + 0.5f) - 0.5f;
        sy = sy < 0.f ? 0.f : sy;
        sx = sx < 0.f ? 0.f : sx;
        const int y0 = (int)sy, x0 = (int)sx;
        const int yp = y0 < h - 1 ? 1 : 0, xp = x0 < w - 1 ? 1 : 0;
        const float ly1 = sy - (float)y0, lx1 = sx - (float)x0;
        const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
        const float4* r0 = src + ((size_t)y0 * w + x0) * C4;
        const float4* r1 = r0 + (size_t)yp * w * C4;
        const float4 v00 = __ldg(r0), v01 = __ldg(r0 + (size_t)xp * C4), v10 = __ldg(r1), v11 = __ldg(r1 + (size_t)xp * C4);
        const float4 top = f4_fma(lx1, v01, f4_scale(lx0, v00));
        const float4 bot = f4_fma(lx1, v11, f4_scale(lx0, v10));
        v = f4_fma(ly1, bot, f4_scale(ly0, top));
      }
      acc = f4_add(acc, v);
    }
    if (relu) acc = make_float4(fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f), fmaxf(acc.z, 0.f), fmaxf(acc.w, 0.f));
    reinterpret_cast<float4*>(out)[e] = acc;
    if (out16) {                                  // fp16 operand copy of the fused map (fp32 residual-stream mode)
      uint2 o;
      *reinterpret_cast<__half2*>(&o.x) = __floats2half2_rn(acc.x, acc.y);
      *reinterpret_cast<__half2*>(&o.y) = __floats2half2_rn(acc.z, acc.w);
      out16[e] = o;
    }
  }
}

// fp16 maps (fp16 camera stem): one thread per (pixel, 8 channels), fp32 arithmetic, 16-byte loads/stores
struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 ld_h8(const uint4* p) {
  const uint4 u = __ldg(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
  F8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__global__ void __launch_bounds__(256) upsample_sum_f16_kernel(UpsTerms T, int n_img, int H, int W, int C8, int relu,
                                                               const float* __restrict__ bias, uint4* __restrict__ out) {
  const int y = blockIdx.y, img = blockIdx.z;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < W * C8; t += gridDim.x * blockDim.x) {
    const int x = t / C8, c8 = t - x * C8;
    const size_t e = ((size_t)img * H + y) * (size_t)W * C8 + t;
    F8 acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = bias ? __ldg(bias + c8 * 8 + i) : 0.f;
#pragma unroll
    for (int k = 0; k < UPS_MAX_TERMS; ++k) {
      if (k >= T.n) break;
      const int h = T.h[k], w = T.w[k];
      const uint4* src = reinterpret_cast<const uint4*>(T.p[k]) + (size_t)img * h * w * C8 + c8;
      if (h == H && w == W) {
        const F8 v = ld_h8(src + ((size_t)y * W + x) * C8);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc.v[i] += v.v[i];
      } else {
        float sy = T.rh[k] * ((float)y + 0.5f) - 0.5f;
        float sx = T.rw[k] * ((float)x + 0.5f) - 0.5f;
        sy = sy < 0.f ? 0.f : sy;
        sx = sx < 0.f ? 0.f : sx;
        const int y0 = (int)sy, x0 = (int)sx;
        const int yp = y0 < h - 1 ? 1 : 0, xp = x0 < w - 1 ? 1 : 0;
        const float ly1 = sy - (float)y0, lx1 = sx - (float)x0;
        const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
        const uint4* r0 = src + ((size_t)y0 * w + x0) * C8;
        const uint4* r1 = r0 + (size_t)yp * w * C8;
        const F8 v00 = ld_h8(r0), v01 = ld_h8(r0 + (size_t)xp * C8), v10 = ld_h8(r1), v11 = ld_h8(r1 + (size_t)xp * C8);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float top = fmaf(lx1, v01.v[i], lx0 * v00.v[i]);
          const float bot = fmaf(lx1, v11.v[i], lx0 * v10.v[i]);
          acc.v[i] += fmaf(ly1, bot, ly0 * top);
        }
      }
    }
    uint4 o;
    __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = acc.v[2 * i], b = acc.v[2 * i + 1];
      if (relu) {
        a = fmaxf(a, 0.f);
        b = fmaxf(b, 0.f);
      }
      o2[i] = __floats2half2_rn(a, b);
    }
    out[e] = o;
  }
}

}  // namespace ls3d

static int upsample_sum_launch(const void* const* terms, const int32_t* term_h, const int32_t* term_w, int32_t n_terms,
                               int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t relu, const float* bias, void* out, void* stream,
                               int vec, void* out16 = nullptr) {
  using namespace ls3d;
  if (n_img <= 0 || H <= 0 || W <= 0) return LS3D_OK;
  if (!terms || !term_h || !term_w || !out || n_terms < 1 || n_terms > UPS_MAX_TERMS || C <= 0 || (C % vec)) return LS3D_ERR_ARG;
  UpsTerms T;
  T.n = n_terms;
  for (int k = 0; k < n_terms; ++k) {
    if (!terms[k] || term_h[k] <= 0 || term_w[k] <= 0 || term_h[k] > H || term_w[k] > W) return LS3D_ERR_ARG;
    T.p[k] = (const float*)terms[k]; T.h[k] = term_h[k]; T.w[k] = term_w[k];
    T.rh[k] = (float)term_h[k] / (float)H;
    T.rw[k] = (float)term_w[k] / (float)W;
  }
  if (H > 65535 || n_img > 65535 || (long long)W * (C / vec) > (1LL << 30)) return LS3D_ERR_ARG;
  const int row_items = W * (C / vec);
  const int threads = row_items >= 256 ? 256 : (row_items >= 128 ? 128 : 64);
  const dim3 grid((unsigned)((row_items + threads - 1) / threads), (unsigned)H, (unsigned)n_img);
  if (vec == 4)
    upsample_sum_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(T, n_img, H, W, C / 4, relu, bias, (float*)out, (uint2*)out16);
  else
    upsample_sum_f16_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(T, n_img, H, W, C / 8, relu, bias, (uint4*)out);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

extern "C" int ls3d_upsample_sum(const float* const* terms, const int32_t* term_h, const int32_t* term_w, int32_t n_terms,
                                 int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t relu, const float* bias, float* out,
                                 void* stream) {
  return upsample_sum_launch((const void* const*)terms, term_h, term_w, n_terms, n_img, H, W, C, relu, bias, out, stream, 4);
}

extern "C" int ls3d_upsample_sum_f16(const void* const* terms, const int32_t* term_h, const int32_t* term_w, int32_t n_terms,
                                     int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t relu, const float* bias, void* out,
                                     void* stream) {
  return upsample_sum_launch(terms, term_h, term_w, n_terms, n_img, H, W, C, relu, bias, out, stream, 8);
}

// fp32 maps + an fp16 operand copy of the result (the input of the next tensor-core convolution), one pass
extern "C" int ls3d_upsample_sum_dual(const float* const* terms, const int32_t* term_h, const int32_t* term_w, int32_t n_terms,
                                      int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t relu, const float* bias, float* out,
                                      void* out16, void* stream) {
  if (!out16 || (((uintptr_t)out16) & 7)) return LS3D_ERR_ARG;
  return upsample_sum_launch((const void* const*)terms, term_h, term_w, n_terms, n_img, H, W, C, relu, bias, out, stream, 4, out16);
}

// fp16 operand copy of an fp32 map (fp32 residual-stream mode: maps produced by a library convolution enter the own kernels)
namespace ls3d {
__global__ void __launch_bounds__(256) cast_f16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long long n4) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(in + e);
    uint2 o;
    *reinterpret_cast<__half2*>(&o.x) = __floats2half2_rn(v.x, v.y);
    *reinterpret_cast<__half2*>(&o.y) = __floats2half2_rn(v.z, v.w);
    out[e] = o;
  }
}
}  // namespace ls3d

extern "C" int ls3d_cast_f16(const float* in, void* out, int64_t n, void* stream) {
  if (n <= 0) return LS3D_OK;
  if (!in || !out || (n & 3) || (((uintptr_t)in) & 15) || (((uintptr_t)out) & 7)) return LS3D_ERR_ARG;
  const long long n4 = n / 4;
  const long long blocks = (n4 + 255) / 256;
  const int grid = (int)(blocks < 148LL * 32 ? blocks : 148LL * 32);
  ls3d::cast_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)in, (uint2*)out, n4);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

// fp32 [npix][3] (channels-last 3-channel images) -> fp16 [npix][8], channels 3..7 zero
namespace ls3d {
__global__ void __launch_bounds__(256) pad3_f16_kernel(const float* __restrict__ in, uint4* __restrict__ out, long long npix) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < npix; e += (long long)gridDim.x * blockDim.x) {
    const float a = __ldg(in + 3 * e), b = __ldg(in + 3 * e + 1), c = __ldg(in + 3 * e + 2);
    uint4 o = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<__half2*>(&o.x) = __floats2half2_rn(a, b);
    *reinterpret_cast<__half2*>(&o.y) = __floats2half2_rn(c, 0.f);
    out[e] = o;
  }
}
}  // namespace ls3d

extern "C" int ls3d_pad3_f16(const float* in, int64_t n_pixels, void* out, void* stream) {
  if (n_pixels <= 0) return LS3D_OK;
  if (!in || !out || (((uintptr_t)out) & 15)) return LS3D_ERR_ARG;
  const long long blocks = (n_pixels + 255) / 256;
  const int grid = (int)(blocks < 148LL * 32 ? blocks : 148LL * 32);
  ls3d::pad3_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, (uint4*)out, n_pixels);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

namespace ls3d {
__global__ void __launch_bounds__(256) cast_f32_kernel(const uint2* __restrict__ in, float4* __restrict__ out, long long n4) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    const uint2 v = __ldg(in + e);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    out[e] = make_float4(a.x, a.y, b.x, b.y);
  }
}
}  // namespace ls3d

extern "C" int ls3d_cast_f32(const void* in, float* out, int64_t n, void* stream) {
  if (n <= 0) return LS3D_OK;
  if (!in || !out || (n & 3) || (((uintptr_t)in) & 7) || (((uintptr_t)out) & 15)) return LS3D_ERR_ARG;
  const long long n4 = n / 4;
  const long long blocks = (n4 + 255) / 256;
  const int grid = (int)(blocks < 148LL * 32 ? blocks : 148LL * 32);
  ls3d::cast_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint2*)in, (float4*)out, n4);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}
