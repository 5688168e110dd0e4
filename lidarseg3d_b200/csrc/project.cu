// Point -> camera projection (the reference does this on the CPU in the data loader).
//
// Replaces the nuScenes branch of LoadPointCloudFromFile (reference det3d/datasets/pipelines/loading.py:373-416:
// lidar -> camera 4x4 transform, pinhole `view_points` (loading.py:67-103), depth > 0 and 1-pixel image margin,
// LATER cameras overwrite earlier ones, cam id starting at 1, -100 where no camera sees the point) followed by the
// rescale to the network input size and the [-1, 1] normalisation of SegImagePreprocess
// (det3d/datasets/pipelines/segpreprocess.py:544-565,654-671): points_cuv = (valid, cam, v, u).
// One thread per point, cameras unrolled in registers; the 4x4 / 3x3 matrices come from constant-cached kernel params.
// Arithmetic follows the numpy original: float64 transforms, float32 storage of (u, v) before rescaling.
#include "common.cuh"
#include "../../include/ls3d.h"

namespace ls3d {

constexpr int PROJ_MAX_CAM = 8;

struct ProjParams {
  double G[12];                 // ref_to_global rows 0..2 (3x4); identity when the caller passes cam_from_lidar directly
  double T[PROJ_MAX_CAM][12];   // cam_from_global (or cam_from_lidar) rows 0..2 (3x4)
  double K[PROJ_MAX_CAM][9];    // intrinsics 3x3
  int ncam, two_stage;
  int img_h, img_w, net_h, net_w;
  float w_ratio, h_ratio;       // float(net) / float(img) evaluated in double on the host, then rounded (numpy semantics)
};

__global__ void project_points_kernel(const float* __restrict__ pts, int ld_p, int xyz_off, int n, ProjParams P,
                                      float* __restrict__ cuv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = pts[(size_t)i * ld_p + xyz_off], y = pts[(size_t)i * ld_p + xyz_off + 1], z = pts[(size_t)i * ld_p + xyz_off + 2];
  if (P.two_stage) {              // lidar -> global first, as the loader does (loading.py:389-392)
    const double gx = P.G[0] * x + P.G[1] * y + P.G[2] * z + P.G[3];
    const double gy = P.G[4] * x + P.G[5] * y + P.G[6] * z + P.G[7];
    const double gz = P.G[8] * x + P.G[9] * y + P.G[10] * z + P.G[11];
    x = gx; y = gy; z = gz;
  }
  float u_sel = -100.f, v_sel = -100.f, cam_sel = -100.f;
  for (int c = 0; c < P.ncam; ++c) {
    const double* T = P.T[c];
    const double* K = P.K[c];
    const double cx = T[0] * x + T[1] * y + T[2] * z + T[3];
    const double cy = T[4] * x + T[5] * y + T[6] * z + T[7];
    const double cz = T[8] * x + T[9] * y + T[10] * z + T[11];
    const double px = K[0] * cx + K[1] * cy + K[2] * cz;
    const double py = K[3] * cx + K[4] * cy + K[5] * cz;
    const double pz = K[6] * cx + K[7] * cy + K[8] * cz;
    const double u = px / pz, v = py / pz;
    if (cz > 0 && u > 1 && u < P.img_w - 1 && v > 1 && v < P.img_h - 1) {
      u_sel = (float)u; v_sel = (float)v; cam_sel = (float)c + 1.f;   // later cameras overwrite (loading.py:407-409)
    }
  }
  // rescale to the resized image - only rows some camera sees, the others keep -100 (img_transforms.py:86-93 is applied to
  // points_cp[cam mask]) - then normalise (segpreprocess.py:654-671); float32 arithmetic like the numpy arrays
  const bool seen = cam_sel >= 1.f;
  const float us = seen ? __fmul_rn(u_sel, P.w_ratio) : u_sel;
  const float vs = seen ? __fmul_rn(v_sel, P.h_ratio) : v_sel;
  float4 o;
  o.x = cam_sel > 0.f ? 1.f : 0.f;
  o.y = P.ncam > 1 ? __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(cam_sel, 1.f), (float)(P.ncam - 1)), 2.f), 1.f) : 0.f;
  o.z = __fsub_rn(__fmul_rn(__fdiv_rn(vs, (float)(P.net_h - 1)), 2.f), 1.f);
  o.w = __fsub_rn(__fmul_rn(__fdiv_rn(us, (float)(P.net_w - 1)), 2.f), 1.f);
  *reinterpret_cast<float4*>(cuv + (size_t)i * 4) = o;
}

}  // namespace ls3d

static int project_launch(const float* points, int32_t ld_p, int32_t xyz_off, int32_t n, const double* ref_to_global,
                          const double* cam_from, const double* intrinsics, int32_t ncam, int32_t img_h, int32_t img_w,
                          int32_t net_h, int32_t net_w, float* points_cuv, void* stream) {
  using namespace ls3d;
  if (n <= 0) return LS3D_OK;
  if (!points || !cam_from || !intrinsics || !points_cuv || ncam < 1 || ncam > PROJ_MAX_CAM) return LS3D_ERR_ARG;
  ProjParams P;
  P.two_stage = ref_to_global ? 1 : 0;
  for (int k = 0; k < 12; ++k) P.G[k] = ref_to_global ? ref_to_global[k] : ((k % 5) == 0 ? 1.0 : 0.0);
  for (int c = 0; c < ncam; ++c) {
    for (int k = 0; k < 12; ++k) P.T[c][k] = cam_from[c * 16 + k];     // 4x4 row-major, rows 0..2
    for (int k = 0; k < 9; ++k) P.K[c][k] = intrinsics[c * 9 + k];
  }
  P.ncam = ncam; P.img_h = img_h; P.img_w = img_w; P.net_h = net_h; P.net_w = net_w;
  P.w_ratio = (float)((double)net_w / (double)img_w);
  P.h_ratio = (float)((double)net_h / (double)img_h);
  project_points_kernel<<<ls3d_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(points, ld_p, xyz_off, n, P, points_cuv);
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}

extern "C" int ls3d_project_points(const float* points, int32_t ld_p, int32_t xyz_off, int32_t n, const double* cam_from_lidar,
                                   const double* intrinsics, int32_t ncam, int32_t img_h, int32_t img_w, int32_t net_h,
                                   int32_t net_w, float* points_cuv, void* stream) {
  return project_launch(points, ld_p, xyz_off, n, nullptr, cam_from_lidar, intrinsics, ncam, img_h, img_w, net_h, net_w,
                        points_cuv, stream);
}

extern "C" int ls3d_project_points_global(const float* points, int32_t ld_p, int32_t xyz_off, int32_t n,
                                          const double* ref_to_global, const double* cams_from_global, const double* intrinsics,
                                          int32_t ncam, int32_t img_h, int32_t img_w, int32_t net_h, int32_t net_w,
                                          float* points_cuv, void* stream) {
  if (!ref_to_global) return LS3D_ERR_ARG;
  return project_launch(points, ld_p, xyz_off, n, ref_to_global, cams_from_global, intrinsics, ncam, img_h, img_w, net_h, net_w,
                        points_cuv, stream);
}
