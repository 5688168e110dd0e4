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

}  // namespace ls3d

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
