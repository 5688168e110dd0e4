// Point -> camera feature sampling (GF-Phase).
//
// Replaces PointSegMSeg3DHead.get_points_image_feature
// (reference det3d/models/point_heads/point_seg_mseg3d_head.py:200-236): a 3-D F.grid_sample over
// (cam, h, w) with mode='bilinear', padding_mode='zeros', align_corners=True, grid order (u, v, cam).
// The 8-tap trilinear form is kept (the camera coordinate lands on an integer slice +-1 ulp, so the
// cross-camera taps carry ~1e-7 weights exactly like the reference).  Feature maps are channels-last
// [B, ncam, H, W, C] so one point's tap is one contiguous 4*C-byte read; half a warp handles a point.
// Rows of invalid points (points_cuv[:,0] != 1) are written as zeros (the reference scatters the valid
// rows into a zero tensor, point_seg_mseg3d_head.py:314-320).
#include <cuda_fp16.h>

#include "common.cuh"
#include "../../include/ls3d.h"

namespace ls3d {

__device__ __forceinline__ int frame_of_row(const int* off, int nf, int i) {
  int f = 0;
  for (int k = 1; k < nf; ++k)
    if (i >= __ldg(off + k)) f = k;
  return f;
}

__device__ __forceinline__ float4 ld_feat4(const float* p) { return ldg_f4(p); }
__device__ __forceinline__ float4 ld_feat4(const __half* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

template <typename T>
__global__ void sample_image_kernel(const T* __restrict__ feat, int ncam, int H, int W, int C,
                                    const float* __restrict__ cuv, int n, const int* __restrict__ point_off, int n_frames,
                                    float* __restrict__ out, int ld_out, int rnd) {
  const int lane16 = threadIdx.x & 15;
  const long long pt = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  if (pt >= n) return;
  const int i = (int)pt;
  const float4 q = ldg_f4(cuv + (size_t)i * 4);  // valid, cam, v (height), u (width)
  const int c4 = C / 4;
  float* dst = out + (size_t)i * ld_out;
  if (q.x != 1.0f) {
    for (int c = lane16; c < c4; c += 16) *reinterpret_cast<float4*>(dst + c * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const int f = frame_of_row(point_off, n_frames, i);
  // grid_sampler_unnormalize, align_corners=True: ((coord + 1) / 2) * (size - 1)
  const float ix = ((q.w + 1.f) / 2.f) * (float)(W - 1);
  const float iy = ((q.z + 1.f) / 2.f) * (float)(H - 1);
  const float iz = ((q.y + 1.f) / 2.f) * (float)(ncam - 1);
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
  const float tx = ix - fx, ty = iy - fy, tz = iz - fz;
  for (int c = lane16; c < c4; c += 16) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dz = 0; dz < 2; ++dz) {
      const int z = z0 + dz;
      const float wz = dz ? tz : 1.f - tz;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int y = y0 + dy;
        const float wy = dy ? ty : 1.f - ty;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int x = x0 + dx;
          const float wx = dx ? tx : 1.f - tx;
          if ((unsigned)z < (unsigned)ncam && (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) {
            const float wgt = wx * wy * wz;
            const float4 v = ld_feat4(feat + ((((size_t)f * ncam + z) * H + y) * W + x) * C + c * 4);
            acc.x = fmaf(wgt, v.x, acc.x); acc.y = fmaf(wgt, v.y, acc.y);
            acc.z = fmaf(wgt, v.z, acc.z); acc.w = fmaf(wgt, v.w, acc.w);
          }
        }
      }
    }
    if (rnd) { acc.x = to_tf32(acc.x); acc.y = to_tf32(acc.y); acc.z = to_tf32(acc.z); acc.w = to_tf32(acc.w); }
    *reinterpret_cast<float4*>(dst + c * 4) = acc;
  }
}

}  // namespace ls3d

extern "C" int ls3d_sample_image_features(const void* feat_nhwc, int32_t feat_fp16, int32_t n_frames, int32_t ncam, int32_t H,
                                          int32_t W, int32_t C, const float* points_cuv, int32_t n, const int32_t* point_off,
                                          float* out, int32_t ld_out, int32_t round_out, void* stream) {
  using namespace ls3d;
  if (n <= 0) return LS3D_OK;
  if (!feat_nhwc || !points_cuv || !point_off || !out || (C & 3) || (ld_out & 3) || ncam < 1) return LS3D_ERR_ARG;
  const long long threads = (long long)n * 16;
  if (feat_fp16)
    sample_image_kernel<__half><<<ls3d_div_up(threads, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)feat_nhwc, ncam, H, W, C, points_cuv, n, point_off, n_frames, out, ld_out, round_out);
  else
    sample_image_kernel<float><<<ls3d_div_up(threads, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float*)feat_nhwc, ncam, H, W, C, points_cuv, n, point_off, n_frames, out, ld_out, round_out);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}
