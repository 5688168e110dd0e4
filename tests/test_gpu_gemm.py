"""Unit parity of the sparse-convolution / Linear engine (ls3d_gather_gemm) in isolation: every launch shape the UNet and
the heads use, against fp64 torch references built (a) straight from the rulebook table and (b) from dense
F.conv3d / F.conv_transpose3d on densified grids.  Tolerance: 2e-5 of the output scale (error-compensated bf16x3 products,
fp32 accumulation - fp32-equivalent)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import sparse as osp

DEV = "cuda"
TOL = 2e-5


def _mods():
    from lidarseg3d_b200 import gemm, ops
    return ops, gemm


def _sites(seed, B, shape, n):
    rng = np.random.default_rng(seed)
    D, H, W = shape
    cells = rng.choice(B * D * H * W, size=n, replace=False)
    return np.stack([cells // (D * H * W), (cells // (H * W)) % D, (cells // W) % H, cells % W], 1).astype(np.int32)


def _clustered_sites(seed, B, shape, n):
    """LiDAR-like occupancy: noisy surfaces, so that SubM offsets have many pairs."""
    rng = np.random.default_rng(seed)
    D, H, W = shape
    out = []
    for b in range(B):
        y = rng.integers(0, H, n)
        x = rng.integers(0, W, n)
        z = np.clip((D / 2 + 0.15 * (y - H / 2) + rng.normal(0, 0.7, n)).astype(np.int64), 0, D - 1)
        c = np.unique(np.stack([np.full(n, b), z, y, x], 1), axis=0)
        out.append(c[rng.permutation(c.shape[0])])
    return np.concatenate(out).astype(np.int32)


def _ref_from_table(x, w_kio, nbr):
    """out[j] = sum_k x[nbr[k, j]] @ W[k] in fp64."""
    K, M = nbr.shape
    out = torch.zeros(M, w_kio.shape[2], dtype=torch.float64)
    xd, wd = x.double(), w_kio.double()
    for k in range(K):
        j = torch.nonzero(nbr[k] >= 0).squeeze(1)
        if j.numel():
            out[j] += xd[nbr[k, j].long()] @ wd[k]
    return out


def _close(out, ref, tol=TOL):
    scale = float(ref.abs().max())
    err = float((out.double().cpu() - ref).abs().max())
    assert err <= tol * max(scale, 1e-30), (err, scale)


def _random_table(g, K, m_out, rows_in, p_hit):
    nbr = torch.randint(0, rows_in, (K, m_out), generator=g, dtype=torch.int32)
    nbr[torch.rand(K, m_out, generator=g) > p_hit] = -1
    return nbr


@pytest.mark.parametrize("m_out", [1, 127, 128, 129, 1000])
@pytest.mark.parametrize("cin,cout", [(32, 32), (64, 32)])
def test_partial_tiles_and_unaligned_tables(m_out, cin, cout):
    """m_out not a multiple of 4 forces the non-bulk rulebook path; 1 / 127 / 129 rows exercise tile edges."""
    _, gemm = _mods()
    g = torch.Generator().manual_seed(m_out * 7 + cin)
    rows_in = 777
    x = torch.randn(rows_in, cin, generator=g)
    w = torch.randn(27, cin, cout, generator=g) / (27 * cin) ** 0.5
    nbr = _random_table(g, 27, m_out, rows_in, 0.4)
    out = gemm.run(x.to(DEV), gemm.PackedWeight(w.to(DEV)), nbr=nbr.to(DEV))
    assert out.shape == (m_out, cout)
    _close(out, _ref_from_table(x, w, nbr))


@pytest.mark.parametrize("cin,cout", [(13, 32), (16, 16), (32, 64), (64, 64), (64, 128), (128, 128), (128, 64), (48, 48),
                                      (256, 128), (128, 256), (96, 112), (32, 176), (64, 192)])
def test_channel_widths(cin, cout):
    """Every (Cin, Cout) of UNetSCN3D at SCALING_RATIO 1-3, Cin = 13 (zero-padded to 16), and n_pad 112 ... 256."""
    _, gemm = _mods()
    g = torch.Generator().manual_seed(cin * 1000 + cout)
    rows_in, m_out = 3000, 2500
    cpad = (cin + 3) // 4 * 4
    x = torch.zeros(rows_in, cpad)
    x[:, :cin] = torch.randn(rows_in, cin, generator=g)
    w = torch.randn(27, cin, cout, generator=g) / (27 * cin) ** 0.5
    nbr = _random_table(g, 27, m_out, rows_in, 0.3)
    s = torch.rand(cout, generator=g) + 0.5
    b = torch.randn(cout, generator=g)
    out = gemm.run(x.to(DEV), gemm.PackedWeight(w.to(DEV)), nbr=nbr.to(DEV), scale=s.to(DEV), shift=b.to(DEV), relu=True)
    ref = torch.relu(_ref_from_table(x[:, :cin], w, nbr) * s.double() + b.double())
    _close(out, ref)


def test_empty_offsets_and_empty_rows():
    """Offsets without any pair are skipped by the kernel; rows without any neighbour must still be written (shift only)."""
    _, gemm = _mods()
    g = torch.Generator().manual_seed(5)
    rows_in, m_out, C = 500, 700, 32
    x = torch.randn(rows_in, C, generator=g)
    w = torch.randn(27, C, C, generator=g) / (27 * C) ** 0.5
    nbr = _random_table(g, 27, m_out, rows_in, 0.5)
    nbr[3:20] = -1                       # 17 empty offsets
    nbr[:, 100:300] = -1                 # rows (and a whole 128-row tile) without neighbours
    b = torch.randn(C, generator=g)
    out = gemm.run(x.to(DEV), gemm.PackedWeight(w.to(DEV)), nbr=nbr.to(DEV), shift=b.to(DEV))
    _close(out, _ref_from_table(x, w, nbr) + b.double())
    assert torch.equal(out[100:300].cpu(), b.expand(200, C))


def test_residual_modes_and_row_mask():
    _, gemm = _mods()
    g = torch.Generator().manual_seed(6)
    n, C = 1500, 64
    x = torch.randn(n, C, generator=g)
    res = torch.randn(n, C, generator=g)
    w = torch.randn(27, C, C, generator=g) / (27 * C) ** 0.5
    nbr = _random_table(g, 27, n, n, 0.3)
    s = torch.rand(C, generator=g) + 0.5
    b = torch.randn(C, generator=g)
    pw = gemm.PackedWeight(w.to(DEV))
    base = _ref_from_table(x, w, nbr) * s.double() + b.double()
    out1 = gemm.run(x.to(DEV), pw, nbr=nbr.to(DEV), scale=s.to(DEV), shift=b.to(DEV), relu=True, res=res.to(DEV), res_mode=1)
    _close(out1, torch.relu(base + res.double()))                         # SparseBasicBlock: relu(bn(conv) + identity)
    out2 = gemm.run(x.to(DEV), pw, nbr=nbr.to(DEV), scale=s.to(DEV), shift=b.to(DEV), relu=True, res=res.to(DEV), res_mode=2)
    _close(out2, torch.relu(base) + res.double())
    mask = torch.zeros(n, 4)
    mask[:, 0] = (torch.rand(n, generator=g) > 0.4).float()
    out3 = gemm.run(x.to(DEV), pw, nbr=nbr.to(DEV), scale=s.to(DEV), shift=b.to(DEV), relu=True, row_mask=mask.to(DEV))
    ref3 = torch.relu(base) * mask[:, :1].double()
    _close(out3, ref3)
    assert float(out3[mask[:, 0] == 0].abs().max()) == 0.0


@pytest.mark.parametrize("C", [16, 32, 48, 64, 128])
def test_concat_and_channel_reduction(C):
    """UR_block (scn_unet.py:163-187): conv_m over [bottom | trans] (2C -> C) + BN + ReLU, then + channel_reduction(cat).
    C = 16 / 48: a 16-column output panel straddles the two source tensors (SCALING_RATIO 1 / 3)."""
    _, gemm = _mods()
    g = torch.Generator().manual_seed(C)
    n = 1100
    bottom, trans = torch.randn(n, C, generator=g), torch.randn(n, C, generator=g)
    w = torch.randn(27, 2 * C, C, generator=g) / (27 * 2 * C) ** 0.5
    nbr = _random_table(g, 27, n, n, 0.3)
    s = torch.rand(C, generator=g) + 0.5
    b = torch.randn(C, generator=g)
    out = gemm.run(bottom.to(DEV), gemm.PackedWeight(w.to(DEV)), x1=trans.to(DEV), nbr=nbr.to(DEV), scale=s.to(DEV),
                   shift=b.to(DEV), relu=True, red=(bottom.to(DEV), trans.to(DEV)))
    cat = torch.cat([bottom, trans], 1)
    ref = torch.relu(_ref_from_table(cat, w, nbr) * s.double() + b.double()) + cat.double().view(n, C, 2).sum(2)
    _close(out, ref)


@pytest.mark.parametrize("geom", [((3, 3, 3), (2, 2, 2), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (0, 1, 1)), ((3, 1, 1), (2, 1, 1), (0, 0, 0))])
def test_strided_and_inverse_conv_vs_dense(geom):
    """SparseConv3d and SparseInverseConv3d compute through the real rulebook kernels vs dense F.conv3d /
    F.conv_transpose3d read at the active sites (SURVEY 8c pin (2))."""
    ops, gemm = _mods()
    ks, st, pd = geom
    B, shape, Ci, Co = 2, (11, 36, 28), 32, 64
    idx = _clustered_sites(3, B, shape, 2500)
    n = idx.shape[0]
    g = torch.Generator().manual_seed(9)
    feats = torch.randn(n, Ci, generator=g)
    K = ks[0] * ks[1] * ks[2]
    w = torch.randn(*ks, Ci, Co, generator=g) / (K * Ci) ** 0.5
    coords = torch.from_numpy(idx).to(DEV)
    grid = ops.grid_from_coords(coords, B, shape, need_perm=True)
    og, oc = ops.grid_strided(coords, B, shape, ks, st, pd)
    down = ops.rulebook_gather(grid, oc, ks, st, pd)
    y = gemm.run(feats.to(DEV), gemm.PackedWeight(w.reshape(K, Ci, Co).to(DEV)), nbr=down)
    dense = torch.zeros(B, Ci, *shape, dtype=torch.float64)
    dense[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]] = feats.double()
    dref = F.conv3d(dense, w.double().permute(4, 3, 0, 1, 2), stride=st, padding=pd)
    o = oc.cpu().long()
    ref = dref[o[:, 0], :, o[:, 1], o[:, 2], o[:, 3]]
    _close(y, ref)
    # every non-zero site of the dense result is an active output site (output set = union of reachable sites)
    assert int((dref.abs().sum(1) > 0).sum()) <= o.shape[0]
    # inverse conv: coarse -> fine with its own weight, restricted to the fine (original) sites
    wi = torch.randn(*ks, Co, Ci, generator=g) / (K * Co) ** 0.5
    up = ops.rulebook_scatter(og, coords, ks, st, pd)
    yc = y.cpu()
    z = gemm.run(y, gemm.PackedWeight(wi.reshape(K, Co, Ci).to(DEV)), nbr=up)
    cd = torch.zeros(B, Co, *og.shape, dtype=torch.float64)
    cd[o[:, 0], :, o[:, 1], o[:, 2], o[:, 3]] = yc.double()
    # conv_transpose3d weight [Cin=Co, Cout=Ci, kz, ky, kx]; fine[i] += coarse[o] . W_inv[k] with i = o*s - p + k
    opad = tuple(shape[a] - ((og.shape[a] - 1) * st[a] - 2 * pd[a] + ks[a]) for a in range(3))
    tref = F.conv_transpose3d(cd, wi.double().permute(3, 4, 0, 1, 2), stride=st, padding=pd, output_padding=opad)
    ref_i = tref[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]]
    _close(z, ref_i)


def test_subm_vs_dense_lidar_like():
    ops, gemm = _mods()
    B, shape, C = 3, (21, 64, 64), 64
    idx = _clustered_sites(1, B, shape, 9000)
    g = torch.Generator().manual_seed(2)
    feats = torch.randn(idx.shape[0], C, generator=g)
    w = torch.randn(3, 3, 3, C, C, generator=g) / (27 * C) ** 0.5
    coords = torch.from_numpy(idx).to(DEV)
    grid = ops.grid_from_coords(coords, B, shape, need_perm=True)
    nbr = ops.rulebook_gather(grid, coords, (3, 3, 3), (1, 1, 1), (1, 1, 1))
    pairs = float((nbr >= 0).sum()) / idx.shape[0]
    assert pairs > 4.0, pairs                                             # LiDAR-like: several pairs per site
    out = gemm.run(feats.to(DEV), gemm.PackedWeight(w.reshape(27, C, C).to(DEV)), nbr=nbr)
    dense = torch.zeros(B, C, *shape, dtype=torch.float64)
    dense[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]] = feats.double()
    dref = F.conv3d(dense, w.double().permute(4, 3, 0, 1, 2), padding=1)
    _close(out, dref[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]])
    assert np.array_equal(nbr.cpu().numpy(), osp.subm_rulebook(idx, shape, 3))


def test_argument_validation():
    """The C ABI rejects shapes the fused epilogue cannot serve instead of computing garbage."""
    _, gemm = _mods()
    n, C = 256, 32
    x = torch.randn(n, C, device=DEV)
    w = gemm.PackedWeight(torch.randn(1, C, 24, device=DEV))
    bad = torch.randn(n, 30, device=DEV)                                  # red_c = 30: not a multiple of 4
    with pytest.raises(RuntimeError):
        gemm.run(x, w, red=(bad, bad))
    ok = torch.randn(n, 32, device=DEV)                                   # red_c = 32 != cout = 24
    with pytest.raises(RuntimeError):
        gemm.run(x, w, red=(ok, ok))


def _plan_views(plan, K, m):
    T = (m + 127) // 128
    hdr = plan.hdr.cpu().numpy()[:T * 32].reshape(T, 32)
    local = plan.local.cpu().numpy().view(np.uint16)[:T * K * 128].reshape(T, K, 128)
    return hdr, local, plan.pool.cpu().numpy()


def test_tile_plan_bit_exact_vs_oracle_single_pass():
    """LiDAR-like rulebook: every tile fits one pass, and the device plan equals oracle/sparse.py::tile_plan bit for bit
    (ascending distinct rows per tile, uint16 local positions)."""
    ops, _ = _mods()
    B, shape = 2, (21, 40, 40)
    idx = _clustered_sites(4, B, shape, 3000)
    idx = idx[np.lexsort((idx[:, 3], idx[:, 2], idx[:, 1], idx[:, 0]))]      # spatially ordered rows (like the UNet's levels)
    coords = torch.from_numpy(np.ascontiguousarray(idx)).to(DEV)
    grid = ops.grid_from_coords(coords, B, shape, need_perm=True)
    nbr = ops.rulebook_gather(grid, coords, (3, 3, 3), (1, 1, 1), (1, 1, 1))
    plan = ops.TilePlan(nbr)
    K, m = nbr.shape
    hdr, local, pool = _plan_views(plan, K, m)
    ref = osp.tile_plan(nbr.cpu().numpy())
    assert (np.diff(ref["stage_off"]) <= 512).all(), "test geometry must fit one pass"
    assert (hdr[:, 0] == 1).all()
    for t in range(hdr.shape[0]):
        kmask, base, cnt = hdr[t, 1], hdr[t, 2], hdr[t, 3]
        rows = ref["stage_rows"][ref["stage_off"][t]:ref["stage_off"][t + 1]]
        assert cnt == rows.size and np.array_equal(pool[base:base + cnt], rows), t
        assert np.array_equal(local[t], ref["local"][t]), t
        active = [k for k in range(K) if (ref["local"][t, k] != 0xFFFF).any()]
        assert kmask == sum(1 << k for k in active) or (not active and kmask == 1)


def test_tile_plan_multi_pass_is_a_partition():
    """A dense random table (each tile touches > 512 distinct rows) must be cut into passes over disjoint offset ranges, each
    staging <= 512 sorted distinct rows through which every pair of its offsets is addressable."""
    ops, _ = _mods()
    g = torch.Generator().manual_seed(8)
    K, m, rows_in = 27, 1000, 50000
    nbr = _random_table(g, K, m, rows_in, 0.6)
    nbr[5] = -1                                          # an empty offset belongs to no pass
    plan = ops.TilePlan(nbr.to(DEV))
    hdr, local, pool = _plan_views(plan, K, m)
    nb = nbr.numpy()
    multi = 0
    for t in range(hdr.shape[0]):
        n_pass = hdr[t, 0]
        assert 1 <= n_pass <= 8
        multi += n_pass > 1
        seen = 0
        r0, r1 = t * 128, min(m, t * 128 + 128)
        for p in range(n_pass):
            kmask, base, cnt = int(hdr[t, 1 + 3 * p]), int(hdr[t, 2 + 3 * p]), int(hdr[t, 3 + 3 * p])
            assert cnt <= 512 and (kmask & seen) == 0 and base % 4 == 0
            seen |= kmask
            rows = pool[base:base + cnt]
            assert (np.diff(rows) > 0).all()
            for k in range(K):
                if not (kmask >> k) & 1:
                    continue
                e = nb[k, r0:r1]
                l = local[t, k, :r1 - r0].astype(np.int64)
                assert ((l == 0xFFFF) == (e < 0)).all()
                ok = e >= 0
                assert (l[ok] < cnt).all() and np.array_equal(rows[l[ok]], e[ok])
            assert (local[t, :, r1 - r0:] == 0xFFFF).all()
        want = sum(1 << k for k in range(K) if (nb[k, r0:r1] >= 0).any())
        assert seen == want
    assert multi > 0


def test_multi_pass_convolution():
    """The convolution kernel over a plan with several passes per tile (dense random table) and 2 K chunks."""
    _, gemm = _mods()
    g = torch.Generator().manual_seed(12)
    rows_in, m_out, cin, cout = 40000, 900, 64, 64
    x = torch.randn(rows_in, cin, generator=g)
    w = torch.randn(27, cin, cout, generator=g) / (27 * cin) ** 0.5
    nbr = _random_table(g, 27, m_out, rows_in, 0.7)
    out = gemm.run(x.to(DEV), gemm.PackedWeight(w.to(DEV)), nbr=nbr.to(DEV))
    _close(out, _ref_from_table(x, w, nbr))


@pytest.mark.parametrize("koff,cin,cout", [(27, 13, 32), (1, 96, 96), (1, 96, 192), (1, 192, 96), (27, 128, 128), (3, 64, 17),
                                           (27, 256, 128)])
def test_pack_bf16x3_c_abi_matches_layout_restatement(koff, cin, cout):
    """ls3d_gemm_pack_bf16x3 (the C-ABI weight packer a reference-side binding calls) writes bit for bit the image the
    documented layout (gemm.PackedWeight._pack_bf16x3, tensor ops) describes - stacked SWIZZLE_64B and wide SWIZZLE_128B."""
    from lidarseg3d_b200 import gemm
    w = torch.randn(koff, cin, cout, generator=torch.Generator().manual_seed(koff * 1000 + cin + cout)).to(DEV)
    pw = gemm.PackedWeight(w)
    ref = pw._pack_bf16x3(w.float().permute(0, 2, 1))
    assert pw.data.numel() == ref.numel()
    assert torch.equal(pw.data.view(torch.int16).reshape(-1), ref.view(torch.int16).reshape(-1))
