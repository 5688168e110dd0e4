// Voxel feature encoders (readers).
//   ls3d_vfe_descriptor : ImprovedMeanVoxelFeatureExtractor.forward / the descriptor part of
//                         TransformerVoxelFeatureExtractor.forward / MeanVoxelFeatureExtractor.forward
//                         (reference det3d/models/readers/voxel_encoder.py:51-58,74-124,202-243)
//   ls3d_vfe_token_attn : the 5-token self-attention core of TransformerEncoderLayerPreNorm
//                         (voxel_encoder.py:149-157; nn.MultiheadAttention over L = 5 slots, no mask)
//   ls3d_vfe_token_max  : final max over the slots (voxel_encoder.py:263)
#include "common.cuh"
#include "../../include/ls3d.h"

namespace ls3d {

constexpr int VFE_MAX_P = 8;
constexpr int VFE_MAX_F = 8;

// mode 0: out[M, F]      = mean                                   (MeanVFE)
// mode 1: out[M, F+8]    = [mean xyz, max xyz, min xyz, mean rest, density, std]  (ImprovedMeanVFE)
// mode 2: out[M*P, 2F+8] = per slot [point features | descriptor]  (TransVFE token input)
__global__ void vfe_descriptor_kernel(const float* __restrict__ voxels, const int* __restrict__ num_points,
                                      int m, int P, int F, int mode, float* __restrict__ out, int ld_out, int rnd) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= m) return;
  float pt[VFE_MAX_P][VFE_MAX_F];
  const float* src = voxels + (size_t)v * P * F;
  for (int s = 0; s < P; ++s)
    for (int c = 0; c < F; ++c) pt[s][c] = src[s * F + c];
  const float nv = (float)num_points[v];
  float mean[VFE_MAX_F];
  for (int c = 0; c < F; ++c) {
    float a = 0.f;
    for (int s = 0; s < P; ++s) a = __fadd_rn(a, pt[s][c]);
    mean[c] = __fdiv_rn(a, nv);
  }
  if (mode == 0) {
    for (int c = 0; c < F; ++c) out[(size_t)v * ld_out + c] = rnd ? to_tf32(mean[c]) : mean[c];
    for (int c = F; c < ld_out; ++c) out[(size_t)v * ld_out + c] = 0.f;
    return;
  }
  float mask[VFE_MAX_P];
  float msum = 0.f;
  for (int s = 0; s < P; ++s) {
    float a = 0.f;
    for (int c = 0; c < F; ++c) a = __fadd_rn(a, pt[s][c]);
    mask[s] = (a != 0.f) ? 1.f : 0.f;  // voxel_encoder.py:87
    msum += mask[s];
  }
  float mx[3], mn[3];
  for (int a = 0; a < 3; ++a) {
    float hi = -INFINITY, lo = INFINITY;
    for (int s = 0; s < P; ++s) {
      const float pen = __fmul_rn(1.f - mask[s], 1e5f);
      hi = fmaxf(hi, __fsub_rn(pt[s][a], pen));
      lo = fminf(lo, __fadd_rn(pt[s][a], pen));
    }
    mx[a] = hi; mn[a] = lo;
  }
  const float density = __fdiv_rn(msum, (float)P);
  float nsum = 0.f;
  for (int s = 0; s < P; ++s) {
    float q = 0.f;
    for (int a = 0; a < 3; ++a) {
      const float d = __fmul_rn(__fsub_rn(pt[s][a], mean[a]), mask[s]);
      q = __fadd_rn(q, __fmul_rn(d, d));
    }
    nsum = __fadd_rn(nsum, __fsqrt_rn(q));
  }
  const float stdv = __fdiv_rn(nsum, nv);
  float desc[VFE_MAX_F + 8];
  int k = 0;
  for (int a = 0; a < 3; ++a) desc[k++] = mean[a];
  for (int a = 0; a < 3; ++a) desc[k++] = mx[a];
  for (int a = 0; a < 3; ++a) desc[k++] = mn[a];
  for (int c = 3; c < F; ++c) desc[k++] = mean[c];
  desc[k++] = density;
  desc[k++] = stdv;
  if (mode == 1) {
    float* dst = out + (size_t)v * ld_out;
    for (int c = 0; c < k; ++c) dst[c] = rnd ? to_tf32(desc[c]) : desc[c];
    for (int c = k; c < ld_out; ++c) dst[c] = 0.f;
  } else {
    for (int s = 0; s < P; ++s) {
      float* dst = out + ((size_t)v * P + s) * ld_out;
      for (int c = 0; c < F; ++c) dst[c] = rnd ? to_tf32(pt[s][c]) : pt[s][c];
      for (int c = 0; c < k; ++c) dst[F + c] = rnd ? to_tf32(desc[c]) : desc[c];
      for (int c = F + k; c < ld_out; ++c) dst[c] = 0.f;
    }
  }
}

// one thread per (voxel, head, query slot): softmax(q.k/sqrt(dh)) v over the P slots of the voxel.
// qkv rows: [q(E) | k(E) | v(E)], row index = voxel * P + slot.
template <int DH>
__global__ void vfe_token_attn_kernel(const float* __restrict__ qkv, int ld_qkv, int m, int P, int H,
                                      float* __restrict__ out, int ld_out, int rnd) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)m * P * H;
  if (t >= total) return;
  const int h = (int)(t % H);
  const long long row = t / H;           // voxel * P + slot
  const long long vox = row / P;
  const int E = H * DH;
  float q[DH];
  const float* qp = qkv + row * ld_qkv + h * DH;
#pragma unroll
  for (int d = 0; d < DH; d += 4) {
    float4 v4 = ldg_f4(qp + d);
    q[d] = v4.x; q[d + 1] = v4.y; q[d + 2] = v4.z; q[d + 3] = v4.w;
  }
  const float scale = rsqrtf((float)DH);
  float sc[VFE_MAX_P];
  float mx = -INFINITY;
  for (int s = 0; s < P; ++s) {
    const float* kp = qkv + (vox * P + s) * ld_qkv + E + h * DH;
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      float4 k4 = ldg_f4(kp + d);
      a = fmaf(q[d], k4.x, a); a = fmaf(q[d + 1], k4.y, a); a = fmaf(q[d + 2], k4.z, a); a = fmaf(q[d + 3], k4.w, a);
    }
    sc[s] = a * scale;
    mx = fmaxf(mx, sc[s]);
  }
  float den = 0.f;
  float o[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) o[d] = 0.f;
  for (int s = 0; s < P; ++s) {
    const float e = __expf(sc[s] - mx);
    den += e;
    const float* vp = qkv + (vox * P + s) * ld_qkv + 2 * E + h * DH;
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      float4 v4 = ldg_f4(vp + d);
      o[d] = fmaf(e, v4.x, o[d]); o[d + 1] = fmaf(e, v4.y, o[d + 1]);
      o[d + 2] = fmaf(e, v4.z, o[d + 2]); o[d + 3] = fmaf(e, v4.w, o[d + 3]);
    }
  }
  const float inv = 1.f / den;
  float* dst = out + row * ld_out + h * DH;
#pragma unroll
  for (int d = 0; d < DH; d += 4)
    *reinterpret_cast<float4*>(dst + d) =
        rnd ? make_float4(to_tf32(o[d] * inv), to_tf32(o[d + 1] * inv), to_tf32(o[d + 2] * inv), to_tf32(o[d + 3] * inv))
            : make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
}

__global__ void vfe_token_max_kernel(const float* __restrict__ x, int ld_x, int m, int P, int E,
                                     float* __restrict__ out, int ld_out, int rnd) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int e4 = E / 4;
  if (t >= (long long)m * e4) return;
  const long long v = t / e4;
  const int c = (int)(t % e4) * 4;
  float4 r = ldg_f4(x + (v * P) * ld_x + c);
  for (int s = 1; s < P; ++s) {
    float4 q = ldg_f4(x + (v * P + s) * ld_x + c);
    r.x = fmaxf(r.x, q.x); r.y = fmaxf(r.y, q.y); r.z = fmaxf(r.z, q.z); r.w = fmaxf(r.w, q.w);
  }
  if (rnd) { r.x = to_tf32(r.x); r.y = to_tf32(r.y); r.z = to_tf32(r.z); r.w = to_tf32(r.w); }
  *reinterpret_cast<float4*>(out + v * ld_out + c) = r;
}

}  // namespace ls3d

extern "C" int ls3d_vfe_descriptor(const float* voxels, const int32_t* num_points, int32_t m, int32_t P,
                                   int32_t F, int32_t mode, float* out, int32_t ld_out, int32_t round_out, void* stream) {
  using namespace ls3d;
  if (m <= 0) return LS3D_OK;
  if (!voxels || !num_points || !out || P < 1 || P > VFE_MAX_P || F < 3 || F > VFE_MAX_F || mode < 0 ||
      mode > 2)
    return LS3D_ERR_ARG;
  const int need = mode == 0 ? F : (mode == 1 ? F + 8 : 2 * F + 8);
  if (ld_out < need) return LS3D_ERR_ARG;
  vfe_descriptor_kernel<<<ls3d_div_up(m, 128), 128, 0, (cudaStream_t)stream>>>(voxels, num_points, m, P, F, mode,
                                                                               out, ld_out, round_out);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

extern "C" int ls3d_vfe_token_attn(const float* qkv, int32_t ld_qkv, int32_t m, int32_t P, int32_t n_head,
                                   int32_t d_head, float* out, int32_t ld_out, int32_t round_out, void* stream) {
  using namespace ls3d;
  if (m <= 0) return LS3D_OK;
  if (!qkv || !out || P < 1 || P > VFE_MAX_P || (ld_qkv & 3) || (ld_out & 3)) return LS3D_ERR_ARG;
  const long long total = (long long)m * P * n_head;
  const int grid = ls3d_div_up(total, 256);
  if (d_head == 16)
    vfe_token_attn_kernel<16><<<grid, 256, 0, (cudaStream_t)stream>>>(qkv, ld_qkv, m, P, n_head, out, ld_out, round_out);
  else if (d_head == 32)
    vfe_token_attn_kernel<32><<<grid, 256, 0, (cudaStream_t)stream>>>(qkv, ld_qkv, m, P, n_head, out, ld_out, round_out);
  else
    return LS3D_ERR_ARG;
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

extern "C" int ls3d_vfe_token_max(const float* x, int32_t ld_x, int32_t m, int32_t P, int32_t E, float* out,
                                  int32_t ld_out, int32_t round_out, void* stream) {
  using namespace ls3d;
  if (m <= 0) return LS3D_OK;
  if (!x || !out || (E & 3) || (ld_x & 3) || (ld_out & 3)) return LS3D_ERR_ARG;
  vfe_token_max_kernel<<<ls3d_div_up((long long)m * (E / 4), 256), 256, 0, (cudaStream_t)stream>>>(x, ld_x, m, P, E,
                                                                                                  out, ld_out, round_out);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}
