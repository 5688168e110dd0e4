"""Camera-side input production of the MSeg3D path, restated on the CPU (numpy): point -> camera projection, image resize,
normalised sampling coordinates.

TEST INFRASTRUCTURE ONLY (checker for tests/, smoke() and bench.py's parity block / CPU arm); nothing under
lidarseg3d_b200/ imports it.

Follows (all in /root/reference):
  * LoadPointCloudFromFile.__call__, SemanticNuscDataset branch      det3d/datasets/pipelines/loading.py:361-416
  * view_points                                                       det3d/datasets/pipelines/loading.py:67-103
  * image_and_points_cp_and_label_resize                              det3d/datasets/pipelines/img_transforms.py:78-99
  * SegImagePreprocess.__call__ (val mode): normalisation of the coordinates   det3d/datasets/pipelines/segpreprocess.py:654-671
  * cv2.resize(image, (W, H)) default INTER_LINEAR on uint8 (third-party: OpenCV, the reference pins no version; restated
    from OpenCV's published fixed-point algorithm: 11-bit coefficients, horizontal pass in int32, vertical pass
    ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2).
Pinned by tests/test_oracle_camera.py against tests/golden/ref_camera_inputs.npz, which oracle/make_golden.py produces by
running the reference's own loader / preprocess classes (and cv2 4.13 itself) inside the build container.
"""
import numpy as np


def view_points(points, view, normalize):
    """loading.py:67-103."""
    viewpad = np.eye(4)
    viewpad[:view.shape[0], :view.shape[1]] = view
    n = points.shape[1]
    points = np.concatenate((points, np.ones((1, n))))
    points = np.dot(viewpad, points)[:3, :]
    if normalize:
        points = points / points[2:3, :].repeat(3, 0).reshape(3, n)
    return points


def project_points_cp(points_xyz, ref_to_global, cams_from_global, intrinsics, im_shape=(900, 1600)):
    """points_cp [N,3] float32 = (cam_id starting at 1, u (width), v (height)), -100 where no camera sees the point
    (loading.py:384-413): lidar -> global -> camera in float64, pinhole divide, depth > 0 and a 1-pixel margin, LATER cameras
    overwrite earlier ones."""
    n = points_xyz.shape[0]
    im_shape = (int(im_shape[0]), int(im_shape[1]))
    pts_uv_all = np.ones([n, 3]).astype(np.float32) * -100
    pts_hom = np.concatenate([points_xyz[:, :3], np.ones([n, 1])], axis=1)
    for cam_id, (cam_from_global, K) in enumerate(zip(cams_from_global, intrinsics)):
        pts_global = np.asarray(ref_to_global).dot(pts_hom.T)
        pts_cam = np.asarray(cam_from_global).dot(pts_global)[:3, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            pts_uv = view_points(pts_cam, np.array(K), normalize=True).T
        mask = (pts_cam[2, :] > 0) & (pts_uv[:, 0] > 1) & (pts_uv[:, 0] < im_shape[1] - 1) & (pts_uv[:, 1] > 1) & \
            (pts_uv[:, 1] < im_shape[0] - 1)
        pts_uv_all[mask, :2] = pts_uv[mask, :2]
        pts_uv_all[mask, 2] = float(cam_id) + 1
    return pts_uv_all[:, [2, 0, 1]]


def points_cuv_from_cp(points_cp, ori_hw, net_hw, ncam):
    """Rescale to the resized image (img_transforms.py:86-93, per camera; all cameras share one size here) and normalise to
    [-1, 1] (segpreprocess.py:654-671): points_cuv [N,4] float32 = (valid, cam, v, u)."""
    cp = np.array(points_cp, dtype=np.float32, copy=True)
    net_hw = (int(net_hw[0]), int(net_hw[1]))          # python ints: numpy integer scalars would promote the maths to fp64
    ncam = int(ncam)
    width_ratio = float(net_hw[1]) / float(ori_hw[1])
    height_ratio = float(net_hw[0]) / float(ori_hw[0])
    # the reference rescales only the rows of points seen by a camera (cam id >= 1); the others keep -100
    seen = cp[:, 0] >= 1
    cp[seen, 1] = cp[seen, 1] * width_ratio
    cp[seen, 2] = cp[seen, 2] * height_ratio
    cuv = np.ones([cp.shape[0], 3]).astype(np.float32) * -100
    if ncam > 1:
        cuv[:, 0] = (cp[:, 0] - 1) / (ncam - 1) * 2 - 1
    else:
        cuv[:, 0] = 0
    cuv[:, 1] = cp[:, 2] / (net_hw[0] - 1) * 2 - 1
    cuv[:, 2] = cp[:, 1] / (net_hw[1] - 1) * 2 - 1
    valid = (cp[:, 0:1] > 0).astype(cuv.dtype)
    return np.concatenate([valid, cuv], axis=1)


def project_points(points_xyz, ref_to_global, cams_from_global, intrinsics, ori_hw, net_hw):
    """points [N,>=3] -> points_cuv [N,4]: the whole camera-coordinate chain of the loader for one frame."""
    cp = project_points_cp(np.asarray(points_xyz), ref_to_global, cams_from_global, intrinsics, im_shape=ori_hw)
    return points_cuv_from_cp(cp, ori_hw, net_hw, len(cams_from_global))


# ------------------------------------------------------------------------------------------------ cv2.resize (uint8, bilinear)
_COEF_BITS = 11
_COEF_SCALE = 1 << _COEF_BITS


def _linear_taps(dst, src, clamp_weights=True):
    """Source index and 11-bit fixed-point weights of every destination coordinate (OpenCV resize, INTER_LINEAR).
    Horizontally OpenCV snaps the weight to (1, 0) when the tap leaves the image; vertically it keeps the fractional weights
    and clamps the two ROW INDICES instead (both taps then read the border row, which rounds differently)."""
    scale = float(src) / float(dst)                      # double
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)     # OpenCV: fx = (float)((dx + 0.5) * scale_x - 0.5)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_weights:
        lo = s < 0
        f[lo], s[lo] = 0.0, 0
        hi = s >= src - 1
        f[hi], s[hi] = 0.0, src - 1
    w1 = np.rint(f * np.float32(_COEF_SCALE)).astype(np.int32)            # saturate_cast<short>(float) rounds to nearest even
    w0 = np.rint((np.float32(1.0) - f) * np.float32(_COEF_SCALE)).astype(np.int32)
    s1 = np.clip(s + 1, 0, src - 1)
    return np.clip(s, 0, src - 1), s1, w0, w1


def resize_bilinear_u8(image, size_wh):
    """cv2.resize(image, (W, H)) for uint8 HxWxC images with the default interpolation (INTER_LINEAR), bit-exact."""
    img = np.asarray(image)
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W = img.shape[:2]
    ow, oh = int(size_wh[0]), int(size_wh[1])
    if (ow, oh) == (W, H):
        return img.copy()
    x0, x1, a0, a1 = _linear_taps(ow, W)
    y0, y1, b0, b1 = _linear_taps(oh, H, clamp_weights=False)
    src = img.astype(np.int32)
    rows = src[:, x0, :] * a0[None, :, None] + src[:, x1, :] * a1[None, :, None]          # horizontal pass, scale 2^11
    r0, r1 = rows[y0], rows[y1]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)
