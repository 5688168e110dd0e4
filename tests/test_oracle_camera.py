"""oracle/camera.py (projection, cv2.resize restatement, sampling-coordinate normalisation) pinned to the outputs of the
REFERENCE's own loader classes (LoadPointCloudFromFile, SegImagePreprocess) recorded in tests/golden/ref_camera_inputs.npz
by oracle/make_golden.py::gen_camera_inputs."""
import hashlib
import os

import numpy as np
import pytest

from lidarseg3d_b200 import synth
from oracle import camera as oc


@pytest.fixture(scope="module")
def cam(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_camera_inputs.npz"))


def test_projection_bit_exact_vs_reference_loader(cam):
    cp = oc.project_points_cp(cam["points"][:, :3], cam["ref_to_global"], cam["cams_from_global"], cam["intrinsics"],
                              im_shape=tuple(cam["img_hw"]))
    assert cp.dtype == np.float32 and np.array_equal(cp, cam["points_cp"])
    cuv = oc.points_cuv_from_cp(cp, tuple(cam["img_hw"]), tuple(cam["net_hw"]), cam["cams_from_global"].shape[0])
    assert cuv.dtype == np.float32 and np.array_equal(cuv, cam["points_cuv"])
    assert 0.3 < cuv[:, 0].mean() < 0.95                       # both seen and unseen points are covered
    assert len(np.unique(cam["points_cp"][:, 0])) == 7         # every camera + "none"


def test_resize_bit_exact_vs_cv2_in_reference_preprocess(cam):
    spec = dict(synth.NUSC)
    raw = synth.camera_images_u8(spec, int(cam["img_seed"]), hw=tuple(cam["img_hw"]))
    assert hashlib.sha256(raw.tobytes()).digest() == cam["raw_sha256"].tobytes(), "synthetic raw images changed"
    nh, nw = (int(v) for v in cam["net_hw"])
    out = np.stack([oc.resize_bilinear_u8(raw[i], (nw, nh)) for i in range(raw.shape[0])])
    assert np.array_equal(out[:, ::40], cam["resized_rows"])
    assert hashlib.sha256(out.tobytes()).digest() == cam["resized_sha256"].tobytes()
    # + the loader's normalisation (already pinned in test_oracle_golden.py) gives the network input
    from oracle import nets as on
    norm = on.image_input_transform(out, synth.IMG_MEAN, synth.IMG_STD)
    ref = cam["images_norm_rows"]
    assert np.abs(norm[:, :, ::80] - ref).max() <= 4e-7 * np.abs(ref).max() + 1e-7


@pytest.mark.parametrize("shape", [(37, 53, 64, 96), (20, 100, 33, 47), (64, 96, 31, 17), (5, 7, 40, 50), (12, 12, 12, 12)])
def test_resize_vs_cv2_when_available(shape):
    cv2 = pytest.importorskip("cv2")
    h, w, oh, ow = shape
    img = np.random.default_rng(h * w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(oc.resize_bilinear_u8(img, (ow, oh)), cv2.resize(img, (ow, oh)))
