// Camera input preparation on the device: uint8 HWC images -> normalised channels-last maps (fp32 or fp16).
//
// Replaces image_input_transform (reference det3d/datasets/pipelines/img_transforms.py:18-29, called per camera in
// datasets/pipelines/segpreprocess.py:621-628) followed by the [H, W, 3] -> [3, H, W] transpose of :637: the reference does
// x / 255 - mean, / std in fp32 numpy on the loader's CPU and ships 4 bytes per value to the GPU; here the uint8 image is
// uploaded (1 byte per value) and the same arithmetic (IEEE fp32 divide, subtract, divide) runs in one HBM-bound pass whose
// output is the channels-last layout the stem convolution reads (logical [n, 3, H, W], physical [n, H, W, 3]).
// Bound: HBM, 3 bytes read + 6 (fp16) / 12 (fp32) bytes written per pixel.
#include <cuda_fp16.h>

#include "common.cuh"
#include "../../include/ls3d.h"

namespace ls3d {

struct Norm3 {
  float mean[3], std[3];
};

__device__ __forceinline__ float norm1(uint32_t v, float mean, float std) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), mean), std);
}

// one thread = 16 pixels = 48 bytes in (3 x 16-byte loads); the channel of byte i of the group is i % 3 (48 % 3 == 0)
template <bool HALF>
__global__ void normalize_u8_kernel(const uint8_t* __restrict__ in, long long n_values, Norm3 p, void* __restrict__ out) {
  const long long n_groups = n_values / 48;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += (long long)gridDim.x * blockDim.x) {
    uint32_t w[12];
    const uint4* src = reinterpret_cast<const uint4*>(in + g * 48);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint4 v = __ldg(src + i);
      w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
    float f[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) f[i] = norm1((w[i >> 2] >> (8 * (i & 3))) & 0xffu, p.mean[i % 3], p.std[i % 3]);
    if (HALF) {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(out) + g * 48);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        uint4 o;
        __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) o2[e] = __floats2half2_rn(f[8 * i + 2 * e], f[8 * i + 2 * e + 1]);
        dst[i] = o;
      }
    } else {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + g * 48);
#pragma unroll
      for (int i = 0; i < 12; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
    }
  }
  // tail (fewer than 48 values): scalar
  for (long long i = n_groups * 48 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_values;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = norm1(in[i], p.mean[i % 3], p.std[i % 3]);
    if (HALF) reinterpret_cast<__half*>(out)[i] = __float2half_rn(v);
    else reinterpret_cast<float*>(out)[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// cv2.resize(image, (W, H)) (uint8, default INTER_LINEAR) fused with the normalisation above.
// Replaces image_and_points_cp_and_label_resize's cv2.resize (reference det3d/datasets/pipelines/img_transforms.py:78-99,
// called per camera from SegImagePreprocess.__call__, segpreprocess.py:544-565) + image_input_transform (:621-628).
// OpenCV's fixed-point algorithm, reproduced bit for bit (oracle/camera.py::resize_bilinear_u8 is the CPU statement of it,
// pinned to cv2 itself): fx = (float)((dx + 0.5) * (src / dst) - 0.5) evaluated in double, 11-bit weights rounded to nearest
// even; horizontally a tap outside the image snaps the weight to (1, 0), vertically the two ROW INDICES are clamped and
// the fractional weights kept; horizontal pass in int32, vertical pass ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2.
// One thread per output pixel (3 channels): 12 byte reads through L1/L2, one 3 / 6 / 12-byte store.  HBM-bound:
// 3 * in_h * in_w bytes read + 3 * out_h * out_w * {1,2,4} written per image.
struct Tap {
  int s0, s1, w0, w1;
};
__device__ __forceinline__ Tap linear_tap(int d, double scale, int src, bool clamp_weights) {
  float f = (float)__dadd_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), -0.5);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (clamp_weights) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  Tap t;
  t.w1 = __float2int_rn(__fmul_rn(f, 2048.f));
  t.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  t.s0 = min(max(s, 0), src - 1);
  t.s1 = min(max(s + 1, 0), src - 1);
  return t;
}

template <int KIND>   // 0: normalised fp32, 1: normalised fp16, 2: resized uint8
__global__ void resize_u8_kernel(const uint8_t* __restrict__ in, int n_img, int in_h, int in_w, int out_h, int out_w,
                                 double scale_x, double scale_y, Norm3 p, void* __restrict__ out) {
  // grid = (column blocks, output rows, images): no 64-bit index division per pixel, the row tap is block-uniform
  const int dy = blockIdx.y, img = blockIdx.z;
  const Tap ty = linear_tap(dy, scale_y, in_h, false);
  for (int dx = blockIdx.x * blockDim.x + threadIdx.x; dx < out_w; dx += gridDim.x * blockDim.x) {
    const size_t i = ((size_t)img * out_h + dy) * out_w + dx;
    const Tap tx = linear_tap(dx, scale_x, in_w, true);
    const uint8_t* r0 = in + ((size_t)img * in_h + ty.s0) * in_w * 3;
    const uint8_t* r1 = in + ((size_t)img * in_h + ty.s1) * in_w * 3;
    uint32_t v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int h0 = (int)__ldg(r0 + tx.s0 * 3 + c) * tx.w0 + (int)__ldg(r0 + tx.s1 * 3 + c) * tx.w1;
      const int h1 = (int)__ldg(r1 + tx.s0 * 3 + c) * tx.w0 + (int)__ldg(r1 + tx.s1 * 3 + c) * tx.w1;
      v[c] = (uint32_t)((((ty.w0 * (h0 >> 4)) >> 16) + ((ty.w1 * (h1 >> 4)) >> 16) + 2) >> 2) & 0xffu;
    }
    if (KIND == 2) {
      uint8_t* o = reinterpret_cast<uint8_t*>(out) + i * 3;
      o[0] = (uint8_t)v[0]; o[1] = (uint8_t)v[1]; o[2] = (uint8_t)v[2];
    } else if (KIND == 1) {
      __half* o = reinterpret_cast<__half*>(out) + i * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) o[c] = __float2half_rn(norm1(v[c], p.mean[c], p.std[c]));
    } else {
      float* o = reinterpret_cast<float*>(out) + i * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) o[c] = norm1(v[c], p.mean[c], p.std[c]);
    }
  }
}

}  // namespace ls3d

extern "C" int ls3d_resize_images_u8(const uint8_t* in, int32_t n_img, int32_t in_h, int32_t in_w, int32_t out_h, int32_t out_w,
                                     const float* mean3, const float* std3, void* out, int32_t out_kind, void* stream) {
  using namespace ls3d;
  if (n_img <= 0) return LS3D_OK;
  if (!in || !out || in_h < 1 || in_w < 1 || out_h < 1 || out_w < 1 || out_kind < 0 || out_kind > 2) return LS3D_ERR_ARG;
  Norm3 p;
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = (out_kind != 2 && mean3) ? mean3[c] : 0.f;
    p.std[c] = (out_kind != 2 && std3) ? std3[c] : 1.f;
    if (!(p.std[c] != 0.f)) return LS3D_ERR_ARG;
  }
  if (out_kind != 2 && (!mean3 || !std3)) return LS3D_ERR_ARG;
  const double sx = (double)in_w / (double)out_w, sy = (double)in_h / (double)out_h;
  if (out_h > 65535 || n_img > 65535) return LS3D_ERR_ARG;
  const dim3 blocks((unsigned)((out_w + 255) / 256), (unsigned)out_h, (unsigned)n_img);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_kind == 0) resize_u8_kernel<0><<<blocks, 256, 0, st>>>(in, n_img, in_h, in_w, out_h, out_w, sx, sy, p, out);
  else if (out_kind == 1) resize_u8_kernel<1><<<blocks, 256, 0, st>>>(in, n_img, in_h, in_w, out_h, out_w, sx, sy, p, out);
  else resize_u8_kernel<2><<<blocks, 256, 0, st>>>(in, n_img, in_h, in_w, out_h, out_w, sx, sy, p, out);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

extern "C" int ls3d_normalize_images_u8(const uint8_t* in, int64_t n_pixels, const float* mean3, const float* std3, void* out,
                                        int32_t out_fp16, void* stream) {
  using namespace ls3d;
  if (n_pixels <= 0) return LS3D_OK;
  if (!in || !out || !mean3 || !std3 || (((uintptr_t)in) & 15) || (((uintptr_t)out) & 15)) return LS3D_ERR_ARG;
  Norm3 p;
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = mean3[c];
    p.std[c] = std3[c];
    if (!(p.std[c] != 0.f)) return LS3D_ERR_ARG;
  }
  const long long n_values = (long long)n_pixels * 3;
  const long long groups = n_values / 48 + 1;
  const int threads = 256;
  long long blocks = (groups + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (out_fp16)
    normalize_u8_kernel<true><<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(in, n_values, p, out);
  else
    normalize_u8_kernel<false><<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(in, n_values, p, out);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}
