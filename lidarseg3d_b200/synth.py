"""Synthetic nuScenes- / SemanticKITTI- / Waymo-shaped inputs.

There are no datasets in the build container or on the GPU box: scans come from a seeded ring-lidar scene model
(ground plane + random vertical walls), cameras from a seeded pinhole rig + ego pose (``calibration``).  The projection
itself is NOT here: the product projects on the GPU (csrc/project.cu) and the checker is oracle/camera.py.
"""
import numpy as np

NUSC = dict(name="nuscenes", beams=32, azimuths=1090, elev=(-30.67, 10.67), feat=5, sensor_h=1.84,
            pc_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], voxel_size=[0.1, 0.1, 0.2], num_class=17,
            ncam=6, cam_yaw=[0, -55, -110, 180, 110, 55], img_hw=(900, 1600), focal=1266.0, net_hw=(640, 960))
KITTI = dict(name="semantickitti", beams=64, azimuths=1900, elev=(-24.8, 2.0), feat=4, sensor_h=1.73,
             pc_range=[-75.2, -75.2, -4.0, 75.2, 75.2, 2.0], voxel_size=[0.1, 0.1, 0.15], num_class=20, ncam=0)
WAYMO = dict(name="waymo", beams=64, azimuths=2500, elev=(-17.6, 2.4), feat=5, sensor_h=2.0,
             pc_range=[-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], voxel_size=[0.1, 0.1, 0.15], num_class=23,
             ncam=5, cam_yaw=[0, 45, -45, 90, -90], img_hw=(1280, 1920), focal=2050.0, net_hw=(640, 960))


def lidar_scan(spec, seed):
    """One frame [N, F] fp32: ranges from a ground plane and random vertical walls at 5-60 m, 2 cm noise."""
    rng = np.random.default_rng(seed)
    el = np.deg2rad(np.linspace(spec["elev"][0], spec["elev"][1], spec["beams"]))
    az = np.linspace(-np.pi, np.pi, spec["azimuths"], endpoint=False)
    E, A = np.meshgrid(el, az, indexing="ij")
    dx, dy, dz = np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)
    h = spec["sensor_h"]
    with np.errstate(divide="ignore", invalid="ignore"):
        r_ground = np.where(dz < -1e-3, -h / dz, np.inf)
    r = r_ground
    nwall = 24
    wd = rng.uniform(5.0, 60.0, nwall)
    wa = rng.uniform(-np.pi, np.pi, nwall)
    ww = rng.uniform(2.0, 15.0, nwall)
    wh = rng.uniform(1.0, 6.0, nwall)
    for d, a, w, hh in zip(wd, wa, ww, wh):
        n = np.array([np.cos(a), np.sin(a)])
        den = dx * n[0] + dy * n[1]
        t = np.where(den > 1e-3, d / np.maximum(den, 1e-3), 1e9)
        px, py, pz = dx * t, dy * t, dz * t + h
        lateral = -px * n[1] + py * n[0]
        hit = (np.abs(lateral) < w / 2) & (pz > 0) & (pz < hh)
        r = np.where(hit & (t < r), t, r)
    ok = np.isfinite(r) & (r < 90.0) & (r > 1.0)
    r = np.where(ok, r, 1.0) + rng.normal(0, 0.02, r.shape)
    x, y, z = dx * r, dy * r, dz * r
    ring = np.repeat(np.arange(spec["beams"])[:, None], spec["azimuths"], 1)
    cols = [x, y, z]
    if spec["feat"] == 5 and spec["name"] == "nuscenes":
        cols += [rng.uniform(0, 255, r.shape), ring.astype(np.float64)]
    elif spec["feat"] == 5:
        cols += [np.tanh(rng.uniform(0, 3, r.shape)), rng.uniform(0, 1, r.shape)]
    else:
        cols += [rng.uniform(0, 1, r.shape)]
    pts = np.stack([c[ok] for c in cols], 1).astype(np.float32)
    return pts


def camera_rig(spec):
    """Per camera: (cam_from_lidar 4x4, intrinsic 3x3).  Camera frame: z forward, x right, y down."""
    H, W = spec["img_hw"]
    K = np.array([[spec["focal"], 0, W / 2.0], [0, spec["focal"], H / 2.0], [0, 0, 1.0]])
    rigs = []
    for yaw in spec["cam_yaw"]:
        a = np.deg2rad(yaw)
        fwd = np.array([np.cos(a), np.sin(a), 0.0])
        right = np.array([np.sin(a), -np.cos(a), 0.0])
        down = np.array([0.0, 0.0, -1.0])
        R = np.stack([right, down, fwd], 0)
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = -R @ np.array([0.0, 0.0, -0.3])
        rigs.append((T, K))
    return rigs


def calibration(spec, seed=0):
    """The calibration record of one frame as the reference's info dict holds it (loading.py:386-388): ref_to_global [4,4]
    (a seeded ego pose), cams_from_global [ncam,4,4], cam_intrinsics [ncam,3,3], plus the raw image size."""
    rng = np.random.default_rng(seed + 104729)
    a = rng.uniform(-np.pi, np.pi)
    G = np.eye(4)
    G[:3, :3] = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    G[:3, 3] = [rng.uniform(-2000, 2000), rng.uniform(-2000, 2000), rng.uniform(-5, 5)]
    Ginv = np.linalg.inv(G)
    rig = camera_rig(spec)
    return dict(ref_to_global=G, cams_from_global=np.stack([T @ Ginv for T, _ in rig]),
                intrinsics=np.stack([K for _, K in rig]), img_hw=tuple(spec["img_hw"]))


def camera_images(spec, seed, hw=None):
    """[ncam, 3, h, w] fp32 'normalised' images (seeded smooth noise); hw defaults to the network input size."""
    rng = np.random.default_rng(seed + 7919)
    h, w = hw or spec["net_hw"]
    low = rng.normal(0, 1, (spec["ncam"], 3, h // 16 + 1, w // 16 + 1)).astype(np.float32)
    img = np.kron(low, np.ones((16, 16), np.float32))[:, :, :h, :w]
    return np.ascontiguousarray(img + rng.normal(0, 0.1, img.shape).astype(np.float32))


# configs/semanticnusc/MSeg3D/semnusc_avgvfe_unetscn3d_hrnetw18_lr1en2_e12.py:19-28 (BGR; the Waymo config reuses them)
IMG_MEAN = [0.40789654, 0.44719302, 0.47026115]
IMG_STD = [0.28863828, 0.27408164, 0.27809835]


def camera_images_u8(spec, seed, hw=None):
    """[ncam, h, w, 3] uint8 camera images as the loader holds them after cv2.resize (seeded smooth noise around mid-grey);
    hw defaults to the network input size."""
    rng = np.random.default_rng(seed + 7919)
    h, w = hw or spec["net_hw"]
    low = rng.normal(0, 1, (spec["ncam"], h // 16 + 1, w // 16 + 1, 3)).astype(np.float32)
    img = np.kron(low, np.ones((1, 16, 16, 1), np.float32))[:, :h, :w]
    img = 112.0 + 64.0 * img + rng.normal(0, 6.0, img.shape).astype(np.float32)
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8))


def grid_shape(spec):
    vs = np.asarray(spec["voxel_size"], np.float32)
    rg = np.asarray(spec["pc_range"], np.float32)
    return np.round((rg[3:] - rg[:3]) / vs).astype(np.int64)
