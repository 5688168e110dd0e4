"""GPU parity tests proper: every C-ABI kernel against the CPU oracle on the same seeded inputs.
Integer / index results are bit-exact; floating point within the tolerances written in each test."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import nets as on
from oracle import sparse as osp
from oracle import voxelize as ov

DEV = "cuda"


def _ops():
    from lidarseg3d_b200 import gemm, ops
    return ops, gemm


# ------------------------------------------------------------------------------------------ voxelize (V1, V2)
@pytest.mark.parametrize("case", ["nusc", "kitti", "capped"])
def test_voxelize_bit_exact_vs_reference_golden(golden_dir, case):
    ops, _ = _ops()
    z = np.load(os.path.join(golden_dir, f"voxelize_{case}.npz"))
    pts = torch.from_numpy(z["points"]).to(DEV)
    r = ops.voxelize(pts, [0, pts.shape[0]], z["voxel_size"].tolist(), z["pc_range"].tolist(), int(z["max_points"]),
                     int(z["max_voxels"]), want_point_map=True)
    assert np.array_equal(r["coordinates"][:, 1:].cpu().numpy(), z["coordinates"])
    assert np.array_equal(r["num_points"].cpu().numpy(), z["num_points"])
    assert np.array_equal(r["voxels"].cpu().numpy(), z["voxels"])
    assert int(r["num_voxels"][0]) == z["voxels"].shape[0]


def test_voxelize_batch_and_edges():
    ops, _ = _ops()
    from lidarseg3d_b200 import synth
    spec = synth.NUSC
    frames = [synth.lidar_scan(spec, s) for s in (0, 1, 2)]
    frames[1] = np.concatenate([frames[1], frames[1][:5000] + np.float32(0.001)])       # many multi-point voxels
    offs = np.cumsum([0] + [f.shape[0] for f in frames]).tolist()
    pts = torch.from_numpy(np.concatenate(frames)).to(DEV)
    r = ops.voxelize(pts, offs, spec["voxel_size"], spec["pc_range"], 5, 300000, want_point_map=True)
    ref = [ov.points_to_voxel(f, spec["voxel_size"], spec["pc_range"], 5, 300000) for f in frames]
    v, c, n, nv, _ = ov.collate_frames([(a, b, cc, f) for (a, b, cc), f in zip(ref, frames)])
    assert np.array_equal(r["num_voxels"].numpy(), nv)
    assert np.array_equal(r["coordinates"].cpu().numpy(), c)
    assert np.array_equal(r["num_points"].cpu().numpy(), n)
    assert np.array_equal(r["voxels"].cpu().numpy(), v)
    # point -> voxel map is consistent with the coordinates
    pm = r["point_voxel"].cpu().numpy()
    cc, valid = ov.voxel_coords_of(np.concatenate(frames), spec["voxel_size"], spec["pc_range"])
    assert np.array_equal(pm >= 0, valid)
    got = r["coordinates"].cpu().numpy()[pm[valid]][:, [3, 2, 1]]
    assert np.array_equal(got, cc[valid])
    # empty frame in the middle, all points out of range
    far = np.full((100, 5), 1000, np.float32)
    r = ops.voxelize(torch.from_numpy(np.concatenate([frames[0], far, frames[2]])).to(DEV),
                     [0, frames[0].shape[0], frames[0].shape[0] + 100, frames[0].shape[0] + 100 + frames[2].shape[0]],
                     spec["voxel_size"], spec["pc_range"], 5, 300000)
    assert r["num_voxels"].tolist() == [ref[0][0].shape[0], 0, ref[2][0].shape[0]]


# ------------------------------------------------------------------------------------------ readers (V3, V3')
def test_vfe_descriptor_vs_reference_golden(ref_modules):
    ops, _ = _ops()
    fx = ref_modules["improved_mean_vfe"]
    out = ops.vfe_descriptor(fx["voxels"].to(DEV), fx["num"].to(DEV), mode=1, ld_out=16)
    torch.testing.assert_close(out[:, :13].cpu(), fx["out"], rtol=1e-5, atol=1e-5)   # fp32, different sum order
    assert float(out[:, 13:].abs().max()) == 0.0
    fx = ref_modules["mean_vfe"]
    out = ops.vfe_descriptor(fx["voxels"].to(DEV), fx["num"].to(DEV), mode=0)
    torch.testing.assert_close(out.cpu(), fx["out"], rtol=1e-6, atol=1e-6)


def test_trans_vfe_vs_reference_golden(ref_modules):
    from lidarseg3d_b200.det3d.readers import TransformerVoxelFeatureExtractor
    from oracle.make_golden import seeded_fill
    fx = ref_modules["trans_vfe"]
    m = TransformerVoxelFeatureExtractor(num_input_features=4, num_compressed_features=16, num_embed=64, num_head=4,
                                         num_layers=3)
    m.load_state_dict(seeded_fill(m.state_dict()))
    m = m.to(DEV).eval()
    out = m(fx["voxels"].to(DEV), fx["num"].to(DEV)).cpu()
    # error-compensated (bf16x3 / 3xTF32) tensor-core GEMMs, fp32 accumulate: 1e-4 of the output scale
    scale = float(fx["out"].abs().max())
    assert float((out - fx["out"]).abs().max()) <= 1e-4 * scale


# ------------------------------------------------------------------------------------------ rulebooks (S1, S2, S4)
def _sites(seed, B, shape, n):
    rng = np.random.default_rng(seed)
    D, H, W = shape
    cells = rng.choice(B * D * H * W, size=n, replace=False)
    cells = np.sort(cells.reshape(B, -1) if False else cells)
    b = cells // (D * H * W)
    order = np.argsort(b, kind="stable")                     # frames contiguous, arbitrary order inside a frame
    cells = cells[order]
    rng2 = np.random.default_rng(seed + 1)
    for f in range(B):
        m = np.nonzero(cells // (D * H * W) == f)[0]
        cells[m] = rng2.permutation(cells[m])
    return np.stack([cells // (D * H * W), (cells // (H * W)) % D, (cells // W) % H, cells % W], 1).astype(np.int32)


@pytest.mark.parametrize("shape,n", [((9, 40, 40), 3000), ((41, 128, 96), 20000)])
def test_rulebooks_bit_exact(shape, n):
    ops, _ = _ops()
    B = 2
    idx = _sites(5, B, shape, n)
    coords = torch.from_numpy(idx).to(DEV)
    g = ops.grid_from_coords(coords, B, shape, need_perm=True)
    assert int(g.total.item()) == n
    nbr = ops.rulebook_gather(g, coords, (3, 3, 3), (1, 1, 1), (1, 1, 1)).cpu().numpy()
    assert np.array_equal(nbr, osp.subm_rulebook(idx, shape, 3))
    for ks, st, pd in [((3, 3, 3), (2, 2, 2), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (0, 1, 1)), ((3, 1, 1), (2, 1, 1), (0, 0, 0))]:
        og, oc = ops.grid_strided(coords, B, shape, ks, st, pd)
        oidx, oshape, nb_ref = osp.strided_rulebook(idx, shape, ks, st, pd)
        assert og.shape == tuple(oshape)
        assert np.array_equal(oc.cpu().numpy(), oidx)                               # ascending linear order
        nb = ops.rulebook_gather(g, oc, ks, st, pd).cpu().numpy()
        assert np.array_equal(nb, nb_ref)
        up = ops.rulebook_scatter(og, coords, ks, st, pd).cpu().numpy()
        assert np.array_equal(up, osp.invert_rulebook(nb_ref, n))
        assert osp.pairs_of(nb) == {(k, i, j) for (k, j, i) in osp.pairs_of(up)}    # same pairs, roles swapped


# ------------------------------------------------------------------------------------------ sparse conv (S3, S5)
def test_sparse_conv_vs_oracle_and_dense():
    ops, gemm = _ops()
    B, shape, C, Co = 2, (9, 32, 32), 32, 64
    idx = _sites(9, B, shape, 4000)
    feats = torch.randn(idx.shape[0], C)
    w = torch.randn(3, 3, 3, C, Co) / (27 * C) ** 0.5
    coords = torch.from_numpy(idx).to(DEV)
    g = ops.grid_from_coords(coords, B, shape, need_perm=True)
    nbr = ops.rulebook_gather(g, coords, (3, 3, 3), (1, 1, 1), (1, 1, 1))
    out = gemm.run(feats.to(DEV), gemm.PackedWeight(w.reshape(27, C, Co).to(DEV)), nbr=nbr).cpu()
    ref = osp.sparse_conv(feats.double(), w.double(), osp.subm_rulebook(idx, shape, 3))
    # error-compensated split products (x_hi.W_hi + x_hi.W_lo + x_lo.W_hi; bf16x3 by default), fp32 accumulate: 2e-5 of the
    # output scale
    assert float((out.double() - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    dense = torch.zeros(B, C, *shape, dtype=torch.float64)
    dense[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]] = feats.double()
    dref = torch.nn.functional.conv3d(dense, w.double().permute(4, 3, 0, 1, 2), padding=1)
    dref = dref[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]]
    assert float((out.double() - dref).abs().max()) <= 2e-5 * float(dref.abs().max())


@pytest.mark.parametrize("m,cin,cout", [(3000, 64, 16), (3000, 32, 48), (3000, 40, 80), (70000, 96, 17), (130, 16, 32)])
def test_dense_gemm_shapes(m, cin, cout):
    """Linear layers of every width used by the heads / TransVFE (n_pad = 16 ... 96), incl. partial tiles."""
    _, gemm = _ops()
    g = torch.Generator().manual_seed(m + cout)
    x = torch.randn(m, cin, generator=g)
    w = torch.randn(cout, cin, generator=g) / cin ** 0.5
    b = torch.randn(cout, generator=g)
    y = gemm.run(x.to(DEV), gemm.PackedWeight.from_linear(w.to(DEV)), shift=b.to(DEV), relu=True).cpu()
    ref = torch.relu(x.double() @ w.double().t() + b.double())
    assert float((y.double() - ref).abs().max()) <= 2e-5 * float(ref.abs().max())


# ------------------------------------------------------------------------------------------ devoxelize (D1, D2)
def test_three_nn_bit_exact_and_interpolate():
    ops, _ = _ops()
    from lidarseg3d_b200 import synth
    spec = synth.NUSC
    frames = [synth.lidar_scan(spec, s) for s in (3, 4)]
    # add far / out-of-range points: exercises the brute-force fallback
    frames[0] = np.concatenate([frames[0], np.array([[80, 80, 10, 0, 0], [-200, 5, 0, 0, 0], [0, 0, 30, 0, 0]], np.float32)])
    offs = np.cumsum([0] + [f.shape[0] for f in frames]).tolist()
    pts = torch.from_numpy(np.concatenate(frames)).to(DEV)
    r = ops.voxelize(pts, offs, spec["voxel_size"], spec["pc_range"], 5, 300000)
    coords = r["coordinates"]
    B = 2
    D, H, W = 41, 1024, 1024
    g = ops.grid_from_coords(coords, B, (D, H, W), need_perm=True)
    bcol = torch.repeat_interleave(torch.arange(B, device=DEV, dtype=torch.float32), torch.tensor(np.diff(offs), device=DEV))
    p4 = torch.cat([bcol[:, None], pts[:, :3]], 1).contiguous()
    poff = torch.tensor(offs, dtype=torch.int32, device=DEV)
    voff = torch.tensor([0] + np.cumsum(r["num_voxels"].numpy()).tolist(), dtype=torch.int32, device=DEV)
    d2, idx = ops.three_nn_grid(p4, g, spec["voxel_size"], spec["pc_range"][:3], poff, voff, coords)
    vs = torch.tensor(spec["voxel_size"]); lo = torch.tensor(spec["pc_range"][:3])
    centers = (coords.cpu()[:, [3, 2, 1]].float() + 0.5) * vs + lo
    feat = torch.randn(coords.shape[0], 32)
    out = ops.three_interpolate(feat.to(DEV), d2, idx).cpu()
    ref_rows = []
    for b in range(B):
        m = coords.cpu()[:, 0] == b
        rd2, ridx = on.three_nn(p4.cpu()[offs[b]:offs[b + 1], 1:4].contiguous(), centers[m].contiguous())
        assert torch.equal(idx.cpu()[offs[b]:offs[b + 1]] - int(voff[b]), ridx)          # bit-exact indices
        assert torch.equal(d2.cpu()[offs[b]:offs[b + 1]], rd2)                           # bit-exact distances
    vcoords = torch.cat([coords.cpu()[:, :1].float(), centers], 1)
    ref = on.three_interpolate_wrap(p4.cpu(), vcoords, feat, B)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------ sampling / SF-Phase
def test_sample_image_features_vs_grid_sample():
    ops, _ = _ops()
    g = torch.Generator().manual_seed(3)
    B, ncam, C, h, w, N = 2, 6, 48, 20, 30, 5000
    img = torch.randn(B, ncam, C, h, w, generator=g)
    cuv = torch.rand(N, 4, generator=g) * 2.2 - 1.1                     # some taps fall outside -> zero padding
    cuv[:, 0] = (torch.rand(N, generator=g) > 0.25).float()
    cuv[:, 1] = torch.randint(0, ncam, (N,), generator=g).float() / (ncam - 1) * 2 - 1
    bidx = (torch.arange(N) >= 2200).float()
    poff = torch.tensor([0, 2200, N], dtype=torch.int32, device=DEV)
    out = ops.sample_image_features(img.permute(0, 1, 3, 4, 2).contiguous().to(DEV), cuv.to(DEV), poff).cpu()
    valid = cuv[:, 0] == 1
    ref = on.sample_image_features(img, cuv[valid], bidx[valid])
    torch.testing.assert_close(out[valid], ref, rtol=1e-5, atol=1e-5)
    assert float(out[~valid].abs().max()) == 0.0


@pytest.mark.parametrize("C,sizes", [(20, [(32, 48), (16, 24), (8, 12), (4, 6)]), (48, [(30, 44), (15, 22)]), (36, [(16, 24), (16, 24), (8, 12)])])
def test_upsample_sum_vs_torch(C, sizes):
    """Fused branch fusion vs the reference's op sequence (F.interpolate bilinear align_corners=False, add, ReLU) in fp32."""
    import torch.nn.functional as F
    ops, _ = _ops()
    g = torch.Generator().manual_seed(5)
    terms = [torch.randn(3, C, h, w, generator=g).to(DEV).contiguous(memory_format=torch.channels_last) for h, w in sizes]
    H, W = sizes[0]
    ref = 0
    for t in terms:
        ref = ref + (t if t.shape[2:] == (H, W) else F.interpolate(t, size=(H, W), mode="bilinear", align_corners=False))
    out = ops.upsample_sum(terms, relu=True)
    assert out.is_contiguous(memory_format=torch.channels_last)
    torch.testing.assert_close(out, torch.relu(ref), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ops.upsample_sum(terms, relu=False), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shape", [(2, 3, 20, 32, 3), (1, 7, 5, 3), (6, 64, 96, 3)])
def test_normalize_images_u8_bit_exact(shape):
    """uint8 HWC -> (x / 255 - mean) / std: fp32 output bit-exact vs the numpy restatement of image_input_transform
    (img_transforms.py:18-29), fp16 output = that result rounded once; both are channels-last views [.., 3, H, W]."""
    from lidarseg3d_b200 import synth
    from oracle import nets as on
    ops, _ = _ops()
    rng = np.random.default_rng(11)
    u8 = rng.integers(0, 256, shape, dtype=np.uint8)
    ref = on.image_input_transform(u8, synth.IMG_MEAN, synth.IMG_STD)
    out = ops.normalize_images_u8(torch.from_numpy(u8).to(DEV), synth.IMG_MEAN, synth.IMG_STD, torch.float32)
    assert tuple(out.shape) == ref.shape
    np.testing.assert_array_equal(out.cpu().numpy(), ref)
    out16 = ops.normalize_images_u8(torch.from_numpy(u8).to(DEV), synth.IMG_MEAN, synth.IMG_STD, torch.float16)
    assert torch.equal(out16.cpu(), torch.from_numpy(ref).half())
    flat = out16.reshape(-1, 3, shape[-3], shape[-2])
    assert flat.is_contiguous(memory_format=torch.channels_last) or flat.shape[0] == 1


@pytest.mark.parametrize("C,sizes", [(24, [(32, 48), (16, 24), (8, 12), (4, 6)]), (48, [(30, 44), (15, 22)])])
def test_upsample_sum_f16_vs_torch(C, sizes):
    """fp16-storage twin of the fused branch fusion: fp32 arithmetic on fp16 maps, one rounding at the store."""
    import torch.nn.functional as F
    ops, _ = _ops()
    g = torch.Generator().manual_seed(6)
    terms = [torch.randn(2, C, h, w, generator=g).half().to(DEV).contiguous(memory_format=torch.channels_last) for h, w in sizes]
    H, W = sizes[0]
    ref = 0
    for t in terms:
        t = t.float()
        ref = ref + (t if t.shape[2:] == (H, W) else F.interpolate(t, size=(H, W), mode="bilinear", align_corners=False))
    out = ops.upsample_sum(terms, relu=True)
    assert out.dtype == torch.float16 and out.is_contiguous(memory_format=torch.channels_last)
    torch.testing.assert_close(out.float(), torch.relu(ref), rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("n,h,w,cin,cout,res,relu", [(1, 16, 8, 16, 16, False, False), (2, 20, 30, 24, 24, True, True),
                                                     (1, 40, 60, 40, 40, True, True), (3, 33, 17, 72, 72, False, True),
                                                     (1, 1, 1, 8, 8, False, True), (2, 7, 129, 24, 40, False, False),
                                                     (1, 160, 240, 24, 24, True, True)])
def test_conv3x3_f16_vs_fp64(n, h, w, cin, cout, res, relu):
    """Fused 3x3 conv + bias (+ residual) (+ ReLU) on fp16 channels-last maps vs an fp64 convolution of the same fp16 operands:
    the only difference allowed is fp32 accumulation order and the final fp16 rounding (2^-11 relative)."""
    import torch.nn.functional as F
    ops, _ = _ops()
    g = torch.Generator().manual_seed(n * 131 + h)
    x = torch.randn(n, cin, h, w, generator=g).half().to(DEV).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    z = torch.randn(n, cout, h, w, generator=g).half().to(DEV).contiguous(memory_format=torch.channels_last) if res else None
    if not ops.conv3x3_f16_supported(cin, cout):
        pytest.skip("weights do not fit shared memory")
    y = ops.conv3x3_f16(x, ops.pack_conv3x3_f16(wt), b, res=z, relu=relu)
    ref = F.conv2d(x.double(), wt.half().double(), b.double(), padding=1)
    if z is not None:
        ref = ref + z.double()
    if relu:
        ref = ref.relu()
    assert y.dtype == torch.float16 and y.is_contiguous(memory_format=torch.channels_last)
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    assert err <= 6e-4, err


@pytest.mark.parametrize("n,h,w,cin,cout,res,relu", [(2, 20, 30, 24, 48, False, True), (1, 40, 60, 72, 24, False, False),
                                                     (3, 33, 17, 64, 256, True, True), (1, 160, 240, 48, 48, False, True),
                                                     (2, 5, 3, 8, 8, False, False)])
def test_conv1x1_f16_vs_fp64(n, h, w, cin, cout, res, relu):
    """The same kernel with a one-tap K list (1x1 convolutions of the fuse layers / Bottlenecks / FCN head), bias optional."""
    import torch.nn.functional as F
    ops, _ = _ops()
    g = torch.Generator().manual_seed(n * 17 + h)
    x = torch.randn(n, cin, h, w, generator=g).half().to(DEV).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV) if relu else None
    z = torch.randn(n, cout, h, w, generator=g).half().to(DEV).contiguous(memory_format=torch.channels_last) if res else None
    if not ops.conv_f16_supported(cin, cout, 1):
        pytest.skip("weights do not fit shared memory")
    y = ops.conv_f16(x, ops.pack_conv_f16(wt), b, res=z, relu=relu, cout=cout, ksize=1)
    ref = F.conv2d(x.double(), wt.half().double(), None if b is None else b.double())
    if z is not None:
        ref = ref + z.double()
    if relu:
        ref = ref.relu()
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    assert err <= 6e-4, err


@pytest.mark.parametrize("n,h,w,cin,cout,res,relu,k", [(2, 20, 30, 24, 24, True, True, 3), (1, 40, 60, 40, 40, False, True, 3),
                                                       (3, 33, 17, 72, 72, True, True, 3), (1, 160, 240, 24, 24, True, True, 3),
                                                       (2, 7, 129, 24, 40, False, False, 3), (1, 40, 60, 72, 24, True, False, 1),
                                                       (1, 1, 1, 8, 8, True, True, 3)])
def test_conv_f16_dual_vs_fp64(n, h, w, cin, cout, res, relu, k):
    """fp32 residual-stream convolution: fp16 operand copy in, fp32 residual in, fp32 map + fp16 operand copy out.  Against an
    fp64 convolution of the same fp16 operands the fp32 map differs by fp32 accumulation order only; the fp16 copy is its RNE
    rounding bit for bit."""
    import torch.nn.functional as F
    ops, _ = _ops()
    g = torch.Generator().manual_seed(n * 977 + h + k)
    x = torch.randn(n, cin, h, w, generator=g).half().to(DEV).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    z = torch.randn(n, cout, h, w, generator=g).to(DEV).contiguous(memory_format=torch.channels_last) if res else None
    if not ops.conv_f16_dual_supported(cin, cout, k):
        pytest.skip("weights do not fit shared memory")
    y32, y16 = ops.conv_f16_dual(x, ops.pack_conv_f16(wt), b, res32=z, relu=relu, cout=cout, ksize=k)
    ref = F.conv2d(x.double(), wt.half().double(), b.double(), padding=k // 2)
    if z is not None:
        ref = ref + z.double()
    if relu:
        ref = ref.relu()
    assert y32.dtype == torch.float32 and y32.is_contiguous(memory_format=torch.channels_last)
    err = float((y32.double() - ref).abs().max() / ref.abs().max())
    assert err <= 2e-6, err
    assert torch.equal(y16, y32.half())


@pytest.mark.parametrize("n,h,w,cin,cout,res,relu,k", [(2, 20, 30, 24, 24, True, True, 3), (1, 40, 60, 40, 40, False, True, 3),
                                                       (1, 160, 240, 24, 24, True, True, 3), (2, 7, 129, 24, 40, False, False, 3),
                                                       (1, 40, 60, 72, 24, True, False, 1), (2, 33, 17, 64, 64, False, True, 1),
                                                       (1, 24, 24, 144, 48, False, False, 1)])
def test_conv_f16_dual_split_weights_vs_fp64(n, h, w, cin, cout, res, relu, k):
    """Split weights [W_hi ; W_lo]: the fp32 weights enter exactly, so against an fp64 convolution of the fp16 ACTIVATIONS with
    the UNROUNDED fp32 weights only fp32 accumulation error remains - and the operand-only variant (no fp32 map) is the RNE
    rounding of the same result."""
    import torch.nn.functional as F
    ops, _ = _ops()
    g = torch.Generator().manual_seed(n * 311 + h + k)
    x = torch.randn(n, cin, h, w, generator=g).half().to(DEV).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    z = torch.randn(n, cout, h, w, generator=g).to(DEV).contiguous(memory_format=torch.channels_last) if res else None
    assert ops.conv_f16_split_supported(cin, cout, k)
    wp = ops.pack_conv_f16_split(wt)
    y32, y16 = ops.conv_f16_dual(x, wp, b, res32=z, relu=relu, cout=cout, ksize=k, split=True)
    ref = F.conv2d(x.double(), wt.double(), b.double(), padding=k // 2)
    if z is not None:
        ref = ref + z.double()
    if relu:
        ref = ref.relu()
    err = float((y32.double() - ref).abs().max() / ref.abs().max())
    assert err <= 2e-6, err
    assert torch.equal(y16, y32.half())
    # the plain (fp16-rounded weights) kernel is measurably further away: the split is doing something
    p32, _ = ops.conv_f16_dual(x, ops.pack_conv_f16(wt), b, res32=z, relu=relu, cout=cout, ksize=k)
    assert float((p32.double() - ref).abs().max() / ref.abs().max()) > 5 * err
    if z is None:
        n32, o16 = ops.conv_f16_dual(x, wp, b, relu=relu, cout=cout, ksize=k, split=True, want32=False)
        assert n32 is None and torch.equal(o16, y16)


@pytest.mark.parametrize("n,h,w,cin,cout,k,stride,res,relu", [
    (2, 32, 48, 8, 64, 3, 2, False, True),          # stem conv1: 3 (-> 8 padded) -> 64, stride 2
    (1, 40, 60, 64, 64, 3, 2, False, True),         # stem conv2
    (2, 20, 30, 24, 40, 3, 2, False, False),        # fuse-layer downsample
    (1, 17, 23, 24, 24, 3, 2, False, True),         # odd sizes: the last output row / column reads the zero padding
    (1, 16, 24, 256, 40, 3, 2, False, True),        # transition from the 256-channel stage: input slices
    (1, 40, 60, 72, 72, 3, 1, True, True),          # 72 channels: the split weight block needs input slices
    (2, 20, 30, 144, 144, 3, 1, True, True),        # 144 channels: output AND input slices
    (1, 40, 60, 64, 256, 1, 1, True, True),         # Bottleneck conv3: two 128-channel output slices
    (1, 40, 60, 256, 64, 1, 1, False, True),        # Bottleneck conv1 of the later blocks
    (1, 24, 36, 144, 48, 1, 1, False, False),       # FCN head branch 1x1
    (3, 9, 7, 48, 24, 1, 1, False, False)])         # conv_seg (17 -> 24 padded classes)
def test_conv_plan_vs_fp64(n, h, w, cin, cout, k, stride, res, relu):
    """ls3d_conv_f16_ex through the launch planner of the camera branch (stride 2 via phase planes, channel slices accumulated
    through the fp32 residual input, split exact weights) against an fp64 convolution of the fp16 activations with the
    UNROUNDED weights."""
    import torch.nn.functional as F
    from lidarseg3d_b200.det3d.img_backbones import ConvPlan
    g = torch.Generator().manual_seed(n * 53 + h + cin + cout)
    x = torch.randn(n, cin, h, w, generator=g).half().to(DEV).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    ho, wo = ((h + 1) // 2, (w + 1) // 2) if stride == 2 else (h, w)
    z = torch.randn(n, cout, ho, wo, generator=g).to(DEV).contiguous(memory_format=torch.channels_last) if res else None
    plan = ConvPlan(wt, b, k, stride, True, True)
    assert plan.ok
    y32, y16 = plan.run(x, res=z, relu=relu)
    ref = F.conv2d(x.double(), wt.double(), b.double(), stride=stride, padding=k // 2)
    if z is not None:
        ref = ref + z.double()
    if relu:
        ref = ref.relu()
    assert y32.shape == ref.shape
    err = float((y32.double() - ref).abs().max() / ref.abs().max())
    assert err <= 3e-6, (err, plan.n_launch)
    assert torch.equal(y16, y32.half())
    op = ConvPlan(wt, b, k, stride, True, False)                      # operand-only plan (fp16 output, fp16 residual)
    if op.ok:
        z16 = None if z is None else z.half()
        n32, o16 = op.run(x, res=z16, relu=relu)
        ref16 = F.conv2d(x.double(), wt.double(), b.double(), stride=stride, padding=k // 2)
        if z16 is not None:
            ref16 = ref16 + z16.double()
        if relu:
            ref16 = ref16.relu()
        assert n32 is None
        assert float((o16.double() - ref16).abs().max() / ref16.abs().max()) <= 6e-4


def test_pad3_f16():
    ops, _ = _ops()
    x = torch.randn(2, 3, 10, 14).to(DEV).contiguous(memory_format=torch.channels_last)
    y = ops.pad3_f16(x)
    assert y.shape == (2, 8, 10, 14) and y.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(y[:, :3], x.half()) and float(y[:, 3:].abs().max()) == 0.0


def test_conv_f16_split_not_supported_shapes():
    ops, _ = _ops()
    assert not ops.conv_f16_split_supported(72, 72, 3)          # doubled weight block does not fit shared memory
    assert not ops.conv_f16_split_supported(64, 256, 1)         # 2 n_pad > 256 accumulator columns
    assert ops.conv_f16_split_supported(24, 24, 3) and ops.conv_f16_split_supported(40, 40, 3)


def test_upsample_sum_dual_and_cast():
    ops, _ = _ops()
    g = torch.Generator().manual_seed(3)
    C, sizes = 24, [(32, 48), (16, 24), (8, 12)]
    terms = [torch.randn(2, C, h, w, generator=g).to(DEV).contiguous(memory_format=torch.channels_last) for h, w in sizes]
    bias = torch.randn(C, generator=g).to(DEV)
    ref = ops.upsample_sum(terms, relu=True, bias=bias)
    o32, o16 = ops.upsample_sum_dual(terms, relu=True, bias=bias)
    assert torch.equal(o32, ref) and torch.equal(o16, ref.half())
    x = torch.randn(3, 20, 9, 7, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    c = ops.cast_f16(x)
    assert c.is_contiguous(memory_format=torch.channels_last) and torch.equal(c, x.half())


@pytest.mark.parametrize("H,dh,L", [(4, 24, 34), (4, 24, 46), (2, 16, 5), (8, 32, 40)])
def test_token_attention_vs_fp64(H, dh, L):
    """Class-token cross attention (context_module.py:320-376) vs a float64 softmax(q K^T) V, frames of uneven size with a
    boundary inside a block; also the same attention through the fused gather-GEMM epilogue (LS3D_EPI_ATTN)."""
    ops, gemm = _ops()
    g = torch.Generator().manual_seed(7)
    m, Fr = 1000, 3
    q = torch.randn(m, H * dh, generator=g).to(DEV)
    K = torch.randn(Fr, H, L, dh, generator=g).to(DEV)
    V = torch.randn(Fr, H, L, dh, generator=g).to(DEV)
    fo = torch.tensor([0, 300, 650], dtype=torch.int32, device=DEV)
    out = ops.token_attention(q, K, V, fo, dh ** -0.5)
    fid = torch.bucketize(torch.arange(m, device=DEV), fo[1:].long(), right=True)
    s = torch.einsum("mhd,mhld->mhl", q.double().view(m, H, dh), K.double()[fid]) * dh ** -0.5
    ref = torch.einsum("mhl,mhld->mhd", s.softmax(-1), V.double()[fid]).reshape(m, H * dh)
    assert float((out.double() - ref).abs().max()) <= 2e-5
    if (H, dh) == (4, 24):
        eye = gemm.PackedWeight(torch.eye(H * dh, device=DEV).unsqueeze(0))
        fused = gemm.run(q, eye, attn=dict(k=K, v=V, frame_off=fo, scale=dh ** -0.5))
        assert float((fused.double() - ref).abs().max()) <= 2e-4


def test_class_embed_and_tokens_vs_oracle(ref_modules):
    ops, _ = _ops()
    from lidarseg3d_b200.det3d.point_heads import PointSegMSeg3DHead
    from oracle.make_golden import HEAD_CFG, seeded_fill
    fx = ref_modules["mseg3d_head"]
    inp = fx["inputs"]
    vf, vl = inp["conv_point_features"], fx["voxel_logits"]
    off = torch.tensor([0, 150, 300], dtype=torch.int32, device=DEV)
    emb = ops.class_embed(vl.to(DEV), vf.to(DEV), off, 2, 300).cpu()                  # [B, ncls, C]
    torch.testing.assert_close(emb, fx["lidar_emb"].squeeze(-1).permute(0, 2, 1), rtol=1e-4, atol=1e-5)
    # class-token memory path vs the oracle's per-layer memory
    m = PointSegMSeg3DHead(class_agnostic=False, num_class=17, model_cfg=HEAD_CFG)
    sd = seeded_fill(m.state_dict())
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    P = m.prep()
    cam = inp["camera_semantic_embeddings"].squeeze(-1).permute(0, 2, 1).contiguous()
    K, V, mem = ops.class_tokens(cam.to(DEV), emb.to(DEV), P["token_params"], 2, 4, 96, want_memory=True)
    geo = torch.randn(500, 64)
    _, mems = on.sffm(sd, "sffm.", geo, inp["camera_semantic_embeddings"], fx["lidar_emb"], inp["points"][:, 0], 2, 4, 2,
                      return_memory=True)
    for l in range(2):
        torch.testing.assert_close(mem[l].cpu(), mems[l].permute(1, 0, 2), rtol=1e-4, atol=1e-4)


def test_mseg3d_head_vs_reference_golden(ref_modules):
    """Whole PointSegMSeg3DHead on the kernels vs the REFERENCE module's output (golden)."""
    ops, _ = _ops()
    from lidarseg3d_b200.det3d.backbones import SparseLevel
    from lidarseg3d_b200.det3d.point_heads import PointSegMSeg3DHead
    from oracle.make_golden import HEAD_CFG, seeded_fill
    fx = ref_modules["mseg3d_head"]
    inp = fx["inputs"]
    m = PointSegMSeg3DHead(class_agnostic=False, num_class=17, model_cfg=HEAD_CFG)
    m.load_state_dict(seeded_fill(m.state_dict()))
    m = m.to(DEV).eval()
    vs, lo = [0.1, 0.1, 0.2], [-10.0, -10.0, -5.0]
    vc = inp["conv_point_coords"]
    idx = torch.cat([vc[:, :1], torch.round((vc[:, [3, 2, 1]] - torch.tensor(lo)[[2, 1, 0]]) / torch.tensor(vs)[[2, 1, 0]] - 0.5)],
                    1).int()
    # duplicates in the fixture's random voxel list are legal for the reference's brute-force 3-NN but not for a voxel
    # grid: keep the fixture only if unique, otherwise fall back to comparing with the oracle on a deduplicated set
    shape = (41, 200, 200)
    coords = idx.to(DEV).contiguous()
    lin = ((idx[:, 0].long() * 41 + idx[:, 1]) * 200 + idx[:, 2]) * 200 + idx[:, 3]
    if lin.unique().numel() != lin.numel():
        pytest.skip("fixture has duplicate voxel cells")
    lv1 = SparseLevel(coords, 2, shape, ops.grid_from_coords(coords, 2, shape, need_perm=True))
    bd = dict(batch_size=2, conv_point_features=inp["conv_point_features"].to(DEV), points=inp["points"].to(DEV),
              points_cuv=inp["points_cuv"].to(DEV), image_features=inp["image_features"].to(DEV),
              camera_semantic_embeddings=inp["camera_semantic_embeddings"].to(DEV), _ls3d_level1=lv1,
              _ls3d_voxel_size=vs, _ls3d_pc_range=lo + [10.0, 10.0, 3.0])
    out = m(bd, return_loss=False)["out_logits"].cpu()
    ref = fx["out_logits"]
    rel = float((out - ref).abs().max() / ref.abs().max())
    agree = float((out.argmax(1) == ref.argmax(1)).float().mean())
    assert rel <= 1e-4 and agree >= 0.999, (rel, agree)          # compensated GEMM chain vs the reference module's fp32 output


@pytest.mark.parametrize("n,nl,L", [(1000, 2, 34), (129, 1, 34), (40000, 6, 34), (777, 3, 46)])
def test_sffm_decoder_fused_vs_fp64(n, nl, L):
    """ls3d_sffm_decoder (all layers of the point stream in one launch; context_module.py:147-171,211-250,320-376) vs a float64
    restatement of forward_post, frames of uneven size with a frame boundary inside a 128-point tile."""
    ops, gemm = _ops()
    from lidarseg3d_b200.det3d.common import linear_pack
    from lidarseg3d_b200.det3d.point_heads import _pack_decoder
    E, H, dh, FFN, Fr = 96, 4, 24, 192, 3
    torch.manual_seed(11 + n)
    lin = lambda i, o: torch.nn.Linear(i, o)
    mods = [dict(q=lin(E, E), o=lin(E, E), l1=lin(E, FFN), l2=lin(FFN, E), n2=torch.nn.LayerNorm(E), n3=torch.nn.LayerNorm(E))
            for _ in range(nl)]
    norm_tgt = torch.nn.LayerNorm(E)
    for md in mods + [dict(n=norm_tgt)]:
        for k, v in md.items():
            if isinstance(v, torch.nn.LayerNorm):
                torch.nn.init.normal_(v.weight, 1.0, 0.2)
                torch.nn.init.normal_(v.bias, 0.0, 0.2)
            v.to(DEV)
    ln = lambda m: (m.weight.detach().float().contiguous(), m.bias.detach().float().contiguous())
    layers = [dict(q=linear_pack(md["q"]), o=linear_pack(md["o"]), l1=linear_pack(md["l1"]), l2=linear_pack(md["l2"]),
                   n2=ln(md["n2"]), n3=ln(md["n3"])) for md in mods]

    class SF:
        d_model, nhead = E, H
        class decoder:
            layers = [type("L", (), dict(linear1=md["l1"]))() for md in mods]
    dec = _pack_decoder(SF, layers, ln(norm_tgt))
    assert dec is not None
    g = torch.Generator().manual_seed(5)
    tgt = torch.randn(n, E, generator=g).to(DEV)
    K = torch.randn(nl, Fr, H, L, dh, generator=g).to(DEV)
    V = torch.randn(nl, Fr, H, L, dh, generator=g).to(DEV)
    fo = torch.tensor([0, int(n * 0.3) + 1, int(n * 0.65) + 3, n], dtype=torch.int32, device=DEV)
    out = ops.sffm_decoder(tgt, dec["w"], dec["vec"], K, V, fo, dh ** -0.5, nl, H, FFN)
    fid = torch.bucketize(torch.arange(n, device=DEV), fo[1:Fr].long(), right=True)
    x = tgt.double()
    F = torch.nn.functional
    for i, md in enumerate(mods):
        d = {k: v.double() for k, v in md.items()}
        q = d["q"](x).view(n, H, dh)
        s = torch.einsum("mhd,mhld->mhl", q, K[i].double()[fid]) * dh ** -0.5
        att = torch.einsum("mhl,mhld->mhd", s.softmax(-1), V[i].double()[fid]).reshape(n, E)
        x = d["n2"](x + d["o"](att))
        x = d["n3"](x + d["l2"](F.relu(d["l1"](x))))
        for v in md.values():
            v.float()
    ref = norm_tgt.double()(x).detach()
    err = float((out.double() - ref).abs().max())
    assert err <= 2e-4, err


def test_spmiddle_resnet_vs_reference_golden(golden_dir):
    """SpMiddleResNetFHD (scn.py:84-177) on the rulebook kernels + gather-GEMM vs the golden produced by the REFERENCE's own
    forward (spconv shim): dense output within 1e-4 of scale, site lists of every level bit-exact."""
    from lidarseg3d_b200.det3d.backbones import SpMiddleResNetFHD
    from oracle.make_golden import unet_fill
    g = torch.load(os.path.join(golden_dir, "ref_backbones.pt"), weights_only=False)["spmiddle"]
    net = SpMiddleResNetFHD(num_input_features=13)
    net.load_state_dict(unet_fill(net.state_dict()))
    net = net.to(DEV).eval()
    with torch.no_grad():
        dense, ms = net(g["voxel_features"].to(DEV), g["coordinates"].to(DEV), 2, list(g["input_shape"]))
    ref = g["dense"]
    assert list(dense.shape) == list(ref.shape)
    assert float((dense.cpu() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    for k, r in g["multi_scale"].items():
        assert torch.equal(ms[k].indices.cpu().int(), r["indices"].int()), k
        assert ms[k].spatial_shape == r["shape"], k
        assert float((ms[k].features.cpu() - r["features"]).abs().max()) <= 1e-4 * float(r["features"].abs().max()), k


@pytest.mark.parametrize("n,h,w,cin,cout,res,relu,stride", [
    (18, 40, 60, 72, 72, True, True, 1),       # the bench's 1/4-resolution branch: 432 tiles, three staged tiles per group
    (9, 40, 52, 72, 72, True, False, 1),       # 189 tiles: the last group is partial
    (18, 20, 30, 144, 144, True, True, 1),     # 1/8-resolution branch: two 72-channel output slices, one tile per item
    (2, 20, 30, 144, 144, False, True, 1),
    (1, 17, 23, 72, 72, False, False, 1),      # partial tiles in both directions
    (18, 80, 120, 40, 72, False, False, 2),    # stride-2 fuse convolutions into the 72- / 144-channel branches (phase planes)
    (18, 40, 60, 72, 144, False, True, 2),
    (3, 41, 59, 24, 72, True, True, 2)])       # odd input size: the last output row / column reads the zero padding
def test_conv_kb_streamed_weights_vs_fp64(n, h, w, cin, cout, res, relu, stride):
    """ls3d_conv_f16_kb (streamed weights, group of staged tiles, block-outer / tile-inner MMA order) through the planner against
    an fp64 convolution; fp32-map plan (dual) and operand-only plan."""
    import torch.nn.functional as F
    from lidarseg3d_b200.det3d.img_backbones import ConvPlan
    g = torch.Generator().manual_seed(n * 31 + h + cin)
    x = torch.randn(n, cin, h, w, generator=g).half().to(DEV).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    ho, wo = ((h + 1) // 2, (w + 1) // 2) if stride == 2 else (h, w)
    z = torch.randn(n, cout, ho, wo, generator=g).to(DEV).contiguous(memory_format=torch.channels_last) if res else None
    plan = ConvPlan(wt, b, 3, stride, True, True, pixels=n * h * w)
    assert plan.ok and plan.kb
    y32, y16 = plan.run(x, res=z, relu=relu)
    ref = F.conv2d(x.double(), wt.double(), b.double(), stride=stride, padding=1)
    if z is not None:
        ref = ref + z.double()
    if relu:
        ref = ref.relu()
    err = float((y32.double() - ref).abs().max() / ref.abs().max())
    assert err <= 3e-6, err
    assert torch.equal(y16, y32.half())
    op = ConvPlan(wt, b, 3, stride, True, False, pixels=n * h * w)
    assert op.ok and op.kb
    z16 = None if z is None else z.half()
    n32, o16 = op.run(x, res=z16, relu=relu)
    ref16 = F.conv2d(x.double(), wt.double(), b.double(), stride=stride, padding=1)
    if z16 is not None:
        ref16 = ref16 + z16.double()
    if relu:
        ref16 = ref16.relu()
    assert n32 is None
    assert float((o16.double() - ref16).abs().max() / ref16.abs().max()) <= 6e-4


def test_frame_offsets_vs_bincount():
    """ls3d_frame_offsets (segment table of a batch-sorted tensor) vs bincount + cumsum; fp32 strided column (points) and int32
    column (voxel coordinates), an empty frame in the middle, an empty tensor."""
    ops, _ = _ops()
    counts = [1000, 0, 37, 4096, 1]
    col = torch.cat([torch.full((c,), b) for b, c in enumerate(counts)])
    ref = torch.tensor([0] + list(np.cumsum(counts)), dtype=torch.int32)
    pts = torch.randn(col.numel(), 6)
    pts[:, 0] = col.float()
    pts = pts.to(DEV)
    assert torch.equal(ops.frame_offsets(pts[:, 0], len(counts)).cpu(), ref)
    coords = torch.zeros(col.numel(), 4, dtype=torch.int32)
    coords[:, 0] = col.int()
    assert torch.equal(ops.frame_offsets(coords.to(DEV)[:, 0], len(counts)).cpu(), ref)
    assert torch.equal(ops.frame_offsets(torch.empty(0, dtype=torch.int32, device=DEV), 3).cpu(), torch.zeros(4, dtype=torch.int32))
