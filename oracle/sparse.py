"""Sparse 3-D convolution oracle: spconv 1.x semantics restated on the CPU (numpy + torch).

TEST INFRASTRUCTURE ONLY: the checker for tests/, __graft_entry__.smoke() and bench.py's CPU baseline / reference arm;
nothing under lidarseg3d_b200/ imports it and the product path has no CPU fallback.

spconv 1.x @ fad3000249d27ca918f2655ff73c41f39b0f3127 is a third-party dependency of the reference
(docs/INSTALL.md:88-99) whose source is NOT under /root/reference; its published algorithm
(rulebook = "indice pairs" per kernel offset, then gather -> mm -> scatter-add per offset) is restated
here, anchored on the reference's call sites det3d/models/backbones/scn_unet.py:15-20,39-46,89-160,
203-210 and SURVEY.md Appendix A.  "parity unpinned" against spconv itself; tests pin it to the dense
ground truth F.conv3d / F.conv_transpose3d.

Conventions: indices [N,4] int32 (b,z,y,x); weight [kz,ky,kx,Cin,Cout]; kernel offset index
k = (kz*KY + ky)*KX + kx; correlation: out[o] += in[o*s - pad + k] . W[k].
"""
import numpy as np
import torch


def _triple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


def _lin(idx, shape):
    D, H, W = shape
    idx = idx.astype(np.int64)
    return ((idx[:, 0] * D + idx[:, 1]) * H + idx[:, 2]) * W + idx[:, 3]


def _lookup(sorted_keys, order, q):
    """row index of key q in the (unsorted) site list, -1 when absent."""
    pos = np.searchsorted(sorted_keys, q)
    pos = np.minimum(pos, sorted_keys.size - 1) if sorted_keys.size else pos
    hit = (sorted_keys[pos] == q) if sorted_keys.size else np.zeros(q.shape, bool)
    return np.where(hit, order[pos], -1)


def out_shape(spatial_shape, ksize, stride, padding, dilation=(1, 1, 1)):
    """floor((D + 2p - d(k-1) - 1)/s) + 1 (Appendix A)."""
    return tuple((spatial_shape[i] + 2 * padding[i] - dilation[i] * (ksize[i] - 1) - 1) // stride[i] + 1
                 for i in range(3))


def subm_rulebook(indices, spatial_shape, ksize=3):
    """SubMConv3d rulebook (get_indice_pairs(subm=True)): output sites == input sites, same order;
    pair (in=i, out=j, k) iff p_i = p_j + k - floor(K/2).  Returns nbr [K^3, N] int32: nbr[k, j] = i or -1."""
    ks = _triple(ksize)
    idx = np.asarray(indices, dtype=np.int64)
    keys = _lin(idx, spatial_shape)
    order = np.argsort(keys, kind="stable")
    skeys = keys[order]
    N = idx.shape[0]
    nbr = np.full((ks[0] * ks[1] * ks[2], N), -1, dtype=np.int32)
    k = 0
    for kz in range(ks[0]):
        for ky in range(ks[1]):
            for kx in range(ks[2]):
                q = idx.copy()
                q[:, 1] += kz - ks[0] // 2
                q[:, 2] += ky - ks[1] // 2
                q[:, 3] += kx - ks[2] // 2
                ok = ((q[:, 1] >= 0) & (q[:, 1] < spatial_shape[0]) & (q[:, 2] >= 0) & (q[:, 2] < spatial_shape[1])
                      & (q[:, 3] >= 0) & (q[:, 3] < spatial_shape[2]))
                r = _lookup(skeys, order, _lin(q, spatial_shape))
                nbr[k] = np.where(ok, r, -1)
                k += 1
    return nbr


def strided_rulebook(indices, spatial_shape, ksize, stride, padding):
    """SparseConv3d rulebook.  o = (i + pad - k)/stride where divisible and in bounds; output active
    set = union, numbered by ascending linear index ((b*D+z)*H+y)*W+x (spconv GPU path).
    Returns (out_indices [M,4] int32, out_spatial_shape, nbr_down [K, M] (input row feeding output j at
    offset k, or -1))."""
    ks, st, pd = _triple(ksize), _triple(stride), _triple(padding)
    oshape = out_shape(spatial_shape, ks, st, pd)
    idx = np.asarray(indices, dtype=np.int64)
    cands = []
    for kz in range(ks[0]):
        for ky in range(ks[1]):
            for kx in range(ks[2]):
                num = idx[:, 1:] + np.asarray(pd)[None] - np.asarray([kz, ky, kx])[None]
                ok = np.all((num % np.asarray(st)[None] == 0) & (num >= 0), axis=1)
                o = num // np.asarray(st)[None]
                ok &= np.all(o < np.asarray(oshape)[None], axis=1)
                cands.append(np.concatenate([idx[ok, :1], o[ok]], 1))
    cand = np.concatenate(cands, 0)
    ukeys = np.unique(_lin(cand, oshape))
    W_, H_, D_ = oshape[2], oshape[1], oshape[0]
    out_idx = np.stack([ukeys // (D_ * H_ * W_), (ukeys // (H_ * W_)) % D_, (ukeys // W_) % H_, ukeys % W_], 1)
    # output-stationary table: input row at o*s - pad + k
    in_keys = _lin(idx, spatial_shape)
    order = np.argsort(in_keys, kind="stable")
    skeys = in_keys[order]
    M = out_idx.shape[0]
    nbr = np.full((ks[0] * ks[1] * ks[2], M), -1, dtype=np.int32)
    k = 0
    for kz in range(ks[0]):
        for ky in range(ks[1]):
            for kx in range(ks[2]):
                q = out_idx.copy()
                q[:, 1] = q[:, 1] * st[0] - pd[0] + kz
                q[:, 2] = q[:, 2] * st[1] - pd[1] + ky
                q[:, 3] = q[:, 3] * st[2] - pd[2] + kx
                ok = ((q[:, 1] >= 0) & (q[:, 1] < spatial_shape[0]) & (q[:, 2] >= 0) & (q[:, 2] < spatial_shape[1])
                      & (q[:, 3] >= 0) & (q[:, 3] < spatial_shape[2]))
                r = _lookup(skeys, order, _lin(np.where(ok[:, None], q, 0), spatial_shape))
                nbr[k] = np.where(ok, r, -1)
                k += 1
    return out_idx.astype(np.int32), oshape, nbr


def invert_rulebook(nbr_down, n_fine):
    """SparseInverseConv3d reuses the strided conv's pairs with in/out swapped (Appendix A):
    fine[i] += coarse[j] . W_inv[k] for every cached pair (i, j, k).  A fine row i appears at most once
    per offset k, so the swapped table is nbr_up[k, i] = j."""
    K, M = nbr_down.shape
    nbr_up = np.full((K, n_fine), -1, dtype=np.int32)
    for k in range(K):
        j = np.nonzero(nbr_down[k] >= 0)[0]
        nbr_up[k, nbr_down[k, j]] = j
    return nbr_up


def pairs_of(nbr):
    """spconv-style pair sets {(k, in, out)} from an output-stationary table (for rulebook parity)."""
    k, j = np.nonzero(nbr >= 0)
    return set(zip(k.tolist(), nbr[k, j].tolist(), j.tolist()))


def sparse_conv(features, weight, nbr):
    """indice_conv: out[j] = sum_k in[nbr[k, j]] @ W[k]; fp32, per-offset gather -> mm -> scatter-add."""
    K = nbr.shape[0]
    w = weight.reshape(K, weight.shape[-2], weight.shape[-1])
    out = torch.zeros(nbr.shape[1], w.shape[2], dtype=features.dtype)
    nbr_t = torch.as_tensor(nbr, dtype=torch.long)
    for k in range(K):
        j = torch.nonzero(nbr_t[k] >= 0).squeeze(1)
        if j.numel() == 0:
            continue
        out.index_add_(0, j, features[nbr_t[k, j]] @ w[k])
    return out


def rulebook_dict(indices, spatial_shape, ksize, stride, padding, subm):
    """O(N*K) python-dict enumeration, independent of the vectorised builders (small cases only).
    Returns (out_indices list, pair set {(k, in, out)})."""
    ks, st, pd = _triple(ksize), _triple(stride), _triple(padding)
    sites = {tuple(int(v) for v in r): i for i, r in enumerate(np.asarray(indices))}
    if subm:
        outs = dict(sites)
        oshape = tuple(spatial_shape)
    else:
        oshape = out_shape(spatial_shape, ks, st, pd)
        cset = set()
        for (b, z, y, x) in sites:
            for kz in range(ks[0]):
                for ky in range(ks[1]):
                    for kx in range(ks[2]):
                        n = (z + pd[0] - kz, y + pd[1] - ky, x + pd[2] - kx)
                        if all(v % s == 0 and v >= 0 for v, s in zip(n, st)):
                            o = tuple(v // s for v, s in zip(n, st))
                            if all(a < m for a, m in zip(o, oshape)):
                                cset.add((b,) + o)
        outs = {c: i for i, c in enumerate(sorted(cset))}
    pairs = set()
    for (b, z, y, x), j in outs.items():
        k = 0
        for kz in range(ks[0]):
            for ky in range(ks[1]):
                for kx in range(ks[2]):
                    if subm:
                        q = (b, z + kz - ks[0] // 2, y + ky - ks[1] // 2, x + kx - ks[2] // 2)
                    else:
                        q = (b, z * st[0] - pd[0] + kz, y * st[1] - pd[1] + ky, x * st[2] - pd[2] + kx)
                    i = sites.get(q)
                    if i is not None:
                        pairs.add((k, i, j))
                    k += 1
    return [c for c, _ in sorted(outs.items(), key=lambda t: t[1])], pairs


# ------------------------------------------------------------------------------------------------------------------
# Tile plan for a gather-once sparse convolution (design prototype for the next gather-GEMM revision):
# output rows are grouped into spatially compact tiles, every tile lists the DISTINCT input rows it needs and addresses them
# through a local index table, so a kernel can stage those rows on chip once and serve all kernel offsets from the stage.
def morton_order(indices):
    """Permutation that sorts sites by (batch, Morton code of (z, y, x)): consecutive rows form compact 3-D blocks."""
    idx = np.asarray(indices).astype(np.int64)

    def part(v):
        v = v.astype(np.uint64) & np.uint64(0x1fffff)
        v = (v | (v << np.uint64(32))) & np.uint64(0x1f00000000ffff)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v
    code = part(idx[:, 3]) | (part(idx[:, 2]) << np.uint64(1)) | (part(idx[:, 1]) << np.uint64(2))
    return np.lexsort((code, idx[:, 0]))


def tile_plan(nbr, out_order=None, tile=128):
    """nbr [K, M_out] (input row per offset and output row, -1 = none) -> plan:
        out_rows [T, tile]      output row handled by each tile slot (-1 = padding), tiles follow ``out_order``
        stage_off [T + 1]       CSR offsets into stage_rows
        stage_rows [S]          distinct input rows of each tile, ascending
        local [T, K, tile]      uint16 position of the pair's input row inside the tile's stage, 0xFFFF = none
    Every pair (k, in, out) of the rulebook appears exactly once."""
    K, M = nbr.shape
    order = np.arange(M) if out_order is None else np.asarray(out_order)
    T = (M + tile - 1) // tile
    out_rows = np.full((T, tile), -1, dtype=np.int64)
    out_rows.reshape(-1)[:M] = order
    stage_off = np.zeros(T + 1, dtype=np.int64)
    stage_rows, local = [], np.full((T, K, tile), 0xFFFF, dtype=np.uint16)
    for t in range(T):
        cols = out_rows[t][out_rows[t] >= 0]
        blk = nbr[:, cols]                                       # [K, n]
        rows = np.unique(blk[blk >= 0])
        assert rows.size < 0xFFFF
        stage_rows.append(rows)
        stage_off[t + 1] = stage_off[t] + rows.size
        pos = np.searchsorted(rows, np.where(blk >= 0, blk, rows[0] if rows.size else 0))
        local[t, :, :cols.size] = np.where(blk >= 0, pos, 0xFFFF).astype(np.uint16)
    return dict(out_rows=out_rows, stage_off=stage_off,
                stage_rows=np.concatenate(stage_rows) if stage_rows else np.zeros(0, np.int64), local=local, tile=tile)


def sparse_conv_tiled(features, weight, plan, n_out):
    """The same convolution as ``sparse_conv`` evaluated tile by tile from a staged copy of the tile's distinct input rows
    (what the planned kernel does: stage once, then one [tile x Cin] x [Cin x Cout] product per kernel offset)."""
    K = plan["local"].shape[1]
    w = weight.reshape(K, weight.shape[-2], weight.shape[-1])
    out = torch.zeros(n_out, w.shape[2], dtype=features.dtype)
    zero = torch.zeros(1, features.shape[1], dtype=features.dtype)
    for t in range(plan["out_rows"].shape[0]):
        rows = plan["stage_rows"][plan["stage_off"][t]:plan["stage_off"][t + 1]]
        stage = torch.cat([features[torch.as_tensor(rows, dtype=torch.long)], zero], 0)      # last row = the "none" row
        loc = torch.as_tensor(plan["local"][t].astype(np.int64))
        loc = torch.where(loc == 0xFFFF, torch.full_like(loc, stage.shape[0] - 1), loc)      # [K, tile]
        acc = torch.zeros(plan["tile"], w.shape[2], dtype=features.dtype)
        for k in range(K):
            acc += stage[loc[k]] @ w[k]
        cols = plan["out_rows"][t]
        ok = cols >= 0
        out[torch.as_tensor(cols[ok], dtype=torch.long)] = acc[torch.as_tensor(np.nonzero(ok)[0], dtype=torch.long)]
    return out
