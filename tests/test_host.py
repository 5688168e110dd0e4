"""Host logic without a GPU: registry / builder / config mirror, C-ABI exports, reference configs instantiate."""
import ctypes
import os
import re

import pytest
import torch

from lidarseg3d_b200 import det3d
from lidarseg3d_b200.det3d import Config, Registry, build_detector, build_from_cfg

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_registry_semantics():
    R = Registry("thing")

    @R.register_module
    class A:
        def __init__(self, x, y=2):
            self.x, self.y = x, y

    assert R.get("A") is A and R.name == "thing"
    with pytest.raises(KeyError):
        R.register_module(A)                                   # duplicate name
    with pytest.raises(TypeError):
        R._register_module(lambda: 0)
    a = build_from_cfg(dict(type="A", x=1), R, default_args=dict(y=5, x=9))
    assert (a.x, a.y) == (1, 5)                                # defaults only fill missing keys
    assert build_from_cfg(dict(type=A, x=3), R).x == 3
    with pytest.raises(KeyError, match="not in the thing registry"):
        build_from_cfg(dict(type="B"), R)
    with pytest.raises(TypeError):
        build_from_cfg(dict(type=3), R)


def test_config_and_alias(tmp_path):
    (tmp_path / "sub_cfg.py").write_text("inner = dict(a=1)\n")
    (tmp_path / "cfg.py").write_text("from addict.addict import Dict\nfrom sub_cfg import inner\n"
                                     "model = dict(type='X', nested=dict(k=[1, 2]))\ninner.update(dict(b=2))\nv = 3\n")
    cfg = Config.fromfile(str(tmp_path / "cfg.py"))
    assert cfg.model.nested.k == [1, 2] and cfg.v == 3 and cfg.inner.b == 2 and cfg["model"]["type"] == "X"
    with pytest.raises(AttributeError):
        cfg.model.missing
    assert "model = dict" in cfg.text
    det3d.install_alias()
    from det3d.models import build_detector as bd                 # reference import names resolve
    from det3d.torchie import Config as C2
    from det3d.utils import Registry as R2
    assert bd is build_detector and C2 is Config and R2 is Registry


def test_cabi_exports_every_declared_symbol():
    from lidarseg3d_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "ls3d.h")).read()
    declared = set(re.findall(r"^int (ls3d_\w+)\(", hdr, flags=re.M))
    assert declared and declared == set(capi.EXPORTS)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert ctypes.sizeof(capi.GemmArgs) % 8 == 0
    # argument errors come back as status codes without touching the device
    n = ctypes.c_int64()
    assert capi.lib().ls3d_voxelize_workspace_bytes(1000, 5, 2, ctypes.byref(n)) == 0 and n.value > 0
    assert capi.lib().ls3d_voxelize_workspace_bytes(1000, 99, 2, ctypes.byref(n)) == -1


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference configs only exist in the build container")
@pytest.mark.parametrize("rel,det,nkeys", [
    ("configs/semanticnusc/MSeg3D/semnusc_avgvfe_unetscn3d_hrnetw18_lr1en2_e12.py", "SegMSeg3DNet", 2251),
    ("configs/semantickitti/SDSeg3D/semkitti_transVFE_unetscn3d_batchloss_e10.py", "SegNet", 291),
    ("configs/semanticwaymo/MSeg3D/semwaymo_avgvfe_unetscn3d_hrnetw18_lr1en2_e12.py", "SegMSeg3DNet", 2251)])
def test_reference_configs_instantiate_unchanged(rel, det, nkeys):
    cfg = Config.fromfile(os.path.join(REF, rel))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # "pretrained weights not found" (tolerated, Appendix D item 15)
        m = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    assert type(m).__name__ == det
    sd = m.state_dict()
    assert len(sd) == nkeys
    assert tuple(sd["backbone.conv_input.0.weight"].shape[:3]) == (3, 3, 3)
    assert "backbone.conv_out.0.weight" in sd and tuple(sd["backbone.conv_out.0.weight"].shape) == (3, 1, 1, 128, 128)
    with pytest.raises(NotImplementedError):
        m(dict(), return_loss=True)


def test_shipped_configs_build():
    for name in ("mseg3d_nuscenes.py", "sdseg3d_semantickitti.py"):
        cfg = Config.fromfile(os.path.join(ROOT, "configs", name))
        m = build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg)
        assert sum(p.numel() for p in m.parameters()) > 1e6


def test_ctypes_signatures_match_header_prototypes():
    """Every `int ls3d_*(...)` prototype of include/ls3d.h is bound in capi with the same number of arguments (a drifted
    ctypes signature would otherwise only show up as a crash on the GPU box)."""
    import re
    from lidarseg3d_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "ls3d.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"\bint\s+(ls3d_\w+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)
    assert len(protos) >= 30
    for name, args in protos:
        n = len([a for a in args.split(",") if a.strip() and a.strip() != "void"])
        if name == "ls3d_gather_gemm":
            assert n == 2                                   # (const ls3d_gemm_args*, void* stream): bound by hand in capi._declare
            continue
        assert name in capi._SIGNATURES, f"{name} is declared in ls3d.h but not bound in capi._SIGNATURES"
        assert len(capi._SIGNATURES[name][0]) == n, (name, n, len(capi._SIGNATURES[name][0]))
    assert set(capi._SIGNATURES) <= {p[0] for p in protos}, set(capi._SIGNATURES) - {p[0] for p in protos}


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no module of the package (or bench.py's product arm imports) may reference it."""
    import re
    pkg = os.path.join(ROOT, "lidarseg3d_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(d, f)


def test_cabi_host_side_queries_and_argument_errors():
    """Entry points that answer on the host (sizes, supported-shape queries) and the argument checks that run before any CUDA
    call: usable without a GPU, same status-code contract (0 ok, -1 bad argument) as include/ls3d.h states."""
    import ctypes

    from lidarseg3d_b200 import capi
    L = capi.lib()
    nbytes, cp, npd = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int32()
    # bf16x3 weight image: one n_pad * 128-byte block per (offset, 32-channel chunk)
    assert L.ls3d_gemm_packed_bytes(27, 13, 32, ctypes.byref(nbytes), ctypes.byref(cp), ctypes.byref(npd)) == 0
    assert (nbytes.value, cp.value, npd.value) == (27 * 1 * 32 * 128, 16, 32)
    assert L.ls3d_gemm_packed_bytes(1, 192, 96, ctypes.byref(nbytes), ctypes.byref(cp), ctypes.byref(npd)) == 0
    assert (nbytes.value, cp.value, npd.value) == (6 * 96 * 128, 192, 96)
    assert L.ls3d_gemm_packed_bytes(1, 96, 300, ctypes.byref(nbytes), None, None) == -1          # cout > 256
    assert L.ls3d_gemm_pack_bf16x3(None, 1, 96, 96, None, None) == -1
    # fused decoder: weight image = q (3 x 12 KB) + out (3 x 12 KB) + linear1 (3 x 24 KB) + linear2 (6 x 12 KB) per layer
    wb, vf = ctypes.c_int64(), ctypes.c_int64()
    assert L.ls3d_sffm_decoder_weight_bytes(6, ctypes.byref(wb), ctypes.byref(vf)) == 0
    assert wb.value == 6 * (12 * 12288 + 3 * 24576) and vf.value == 6 * 864 + 192
    assert L.ls3d_sffm_decoder_weight_bytes(9, ctypes.byref(wb), ctypes.byref(vf)) == -1          # > 8 layers
    fake = 1 << 20                                                                             # never dereferenced: the shape check fails first
    assert L.ls3d_sffm_decoder(fake, 96, 10, fake, fake, fake, fake, fake, 1, 34, 6, 4, 64, 192, 1, 0.2, 1e-5, fake, 96, None) == -1
    assert L.ls3d_sffm_decoder(fake, 96, 10, fake, fake, fake, fake, fake, 1, 34, 6, 8, 96, 192, 1, 0.2, 1e-5, fake, 96, None) == -1
    assert L.ls3d_sffm_decoder(None, 96, 0, None, None, None, None, None, 1, 34, 6, 4, 96, 192, 1, 0.2, 1e-5, None, 96, None) == 0   # n = 0
    # streamed-weight convolution: shapes it serves / refuses
    ok = ctypes.c_int32()
    assert L.ls3d_conv_f16_kb_supported(72, 72, 3, 1, 1, 1, 18 * 40 * 60, ctypes.byref(ok)) == 0 and ok.value == 1
    assert L.ls3d_conv_f16_kb_supported(144, 72, 3, 1, 1, 1, 18 * 20 * 30, ctypes.byref(ok)) == 0 and ok.value == 1
    assert L.ls3d_conv_f16_kb_supported(72, 72, 3, 2, 1, 1, 18 * 20 * 30, ctypes.byref(ok)) == 0 and ok.value == 1
    assert L.ls3d_conv_f16_kb_supported(72, 144, 3, 1, 1, 1, 18 * 40 * 60, ctypes.byref(ok)) == 0 and ok.value == 0   # slice > 128 ch
    assert L.ls3d_conv_f16_kb_supported(72, 72, 5, 1, 1, 1, 1000, ctypes.byref(ok)) == -1
    assert L.ls3d_conv_f16_kb_supported(70, 72, 3, 1, 1, 1, 1000, ctypes.byref(ok)) == -1                              # cin % 8
    # segment table, tile plan sizes
    assert L.ls3d_frame_offsets(None, 1, 6, 100, 3, fake, None) == -1
    assert L.ls3d_frame_offsets(fake, 1, 6, 100, 0, fake, None) == -1
    hb, lb, pb = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    assert L.ls3d_tile_plan_bytes(27, 1000, ctypes.byref(hb), ctypes.byref(lb), ctypes.byref(pb)) == 0
    assert hb.value == 8 * 32 * 4 + 16 and lb.value == 8 * 27 * 128 * 2 + 16
    assert L.ls3d_tile_plan_bytes(28, 1000, ctypes.byref(hb), ctypes.byref(lb), ctypes.byref(pb)) == -1               # > 27 offsets


def test_conv_plan_routes_weight_heavy_convolutions_to_the_streamed_weight_kernel(monkeypatch):
    """ConvPlan (det3d/img_backbones.py) at the bench's HRNet-w18 shapes: the 72- / 144-channel 3x3 convolutions and the stride-2
    fuse convolutions into those branches go to ls3d_conv_f16_kb (whole input range per launch, <= 80-channel output slices);
    the high-resolution small-channel convolutions and every 1x1 stay on the resident-weight kernel.  Host logic only."""
    import torch

    from lidarseg3d_b200 import ops
    from lidarseg3d_b200.det3d.img_backbones import ConvPlan
    monkeypatch.setattr(ops, "pack_conv_ex", lambda w, stride, split: torch.zeros(1))      # geometry only, no device
    n = 18

    def plan(co, ci, k, st, f32, h, w):
        return ConvPlan(torch.zeros(co, ci, k, k), torch.zeros(co), k, st, exact=True, fp32out=f32, pixels=n * h * w)

    for co, ci, st, f32, h, w, slices in [(72, 72, 1, True, 40, 60, 1), (72, 72, 1, False, 40, 60, 1), (144, 144, 1, True, 20, 30, 2),
                                          (144, 144, 1, False, 20, 30, 2), (72, 40, 2, True, 80, 120, 1), (144, 72, 2, True, 40, 60, 2),
                                          (144, 24, 2, True, 40, 60, 2)]:
        p = plan(co, ci, 3, st, f32, h, w)
        assert p.ok and p.kb and p.ci == ci and len(p.passes) == slices and p.cs <= 80, (co, ci, st, f32)
        assert [pp[1] for pp in p.passes] == [0] * slices                      # every slice reads the whole input range
    for co, ci, k, st, f32, h, w in [(24, 24, 3, 1, True, 160, 240), (40, 40, 3, 1, False, 80, 120), (40, 24, 3, 2, True, 160, 240),
                                     (256, 64, 1, 1, True, 160, 240), (64, 256, 1, 1, False, 160, 240), (24, 256, 3, 1, True, 160, 240),
                                     (64, 64, 3, 2, True, 320, 480)]:
        p = plan(co, ci, k, st, f32, h, w)
        assert p.ok and not p.kb, (co, ci, k, st)
