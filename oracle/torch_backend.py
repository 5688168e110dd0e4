"""TEST INFRASTRUCTURE ONLY - the sparse-convolution / 3-NN parts of the oracle restated in pure torch so that the whole
reference-algorithm forward (oracle/nets.py) can run on a CUDA device as the "spconv-1.x-style GPU baseline" of SURVEY.md
section 8(d): per kernel offset gather (index_select) -> mm -> scatter-add (index_add_), separate BatchNorm / ReLU, fp32 with
TF32 off, plain per-frame python loops in the heads.  It is the denominator of the north-star ">= 10x the reference
spconv-GPU forward" target - a reported baseline, not product code; nothing under lidarseg3d_b200/ imports it.

Same interface as oracle/sparse.py (which stays the numpy ground truth these functions are tested against on the CPU):
tables ``nbr[k, j]`` = input row feeding output row j at kernel offset k, -1 = none (spconv 1.x semantics).
"""
import torch


def _triple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


def as_indices(voxel_coords):
    return voxel_coords.long()


def out_shape(spatial_shape, ksize, stride, padding):
    return tuple((spatial_shape[i] + 2 * padding[i] - (ksize[i] - 1) - 1) // stride[i] + 1 for i in range(3))


def _lin(idx, shape):
    D, H, W = shape
    return ((idx[:, 0] * D + idx[:, 1]) * H + idx[:, 2]) * W + idx[:, 3]


def _lookup(skeys, order, q):
    if skeys.numel() == 0:
        return torch.full_like(q, -1)
    pos = torch.searchsorted(skeys, q).clamp_(max=skeys.numel() - 1)
    return torch.where(skeys[pos] == q, order[pos], torch.full_like(q, -1))


def _offsets(ks, device):
    kz, ky, kx = torch.meshgrid(torch.arange(ks[0], device=device), torch.arange(ks[1], device=device),
                                torch.arange(ks[2], device=device), indexing="ij")
    return torch.stack([kz.reshape(-1), ky.reshape(-1), kx.reshape(-1)], 1)          # [K, 3], index (kz*K+ky)*K+kx


def _table(sites_idx, sites_shape, query_zyx, query_b):
    """rows of ``sites`` found at the [K, M, 3] query positions (batch query_b [M]) or -1."""
    keys = _lin(sites_idx, sites_shape)
    skeys, order = torch.sort(keys, stable=True)
    sh = torch.tensor(sites_shape, device=sites_idx.device)
    ok = ((query_zyx >= 0) & (query_zyx < sh)).all(-1)                                 # [K, M]
    q = torch.where(ok[..., None], query_zyx, torch.zeros_like(query_zyx))
    D, H, W = sites_shape
    qk = ((query_b[None, :] * D + q[..., 0]) * H + q[..., 1]) * W + q[..., 2]
    r = _lookup(skeys, order, qk.reshape(-1)).reshape(qk.shape)
    return torch.where(ok, r, torch.full_like(r, -1))


def subm_rulebook(indices, spatial_shape, ksize=3):
    ks = _triple(ksize)
    idx = indices.long()
    off = _offsets(ks, idx.device) - torch.tensor([k // 2 for k in ks], device=idx.device)
    return _table(idx, spatial_shape, idx[None, :, 1:] + off[:, None, :], idx[:, 0])


def strided_rulebook(indices, spatial_shape, ksize, stride, padding):
    ks, st, pd = _triple(ksize), _triple(stride), _triple(padding)
    oshape = out_shape(spatial_shape, ks, st, pd)
    idx = indices.long()
    dev = idx.device
    off = _offsets(ks, dev)
    st_t, pd_t, os_t = (torch.tensor(v, device=dev) for v in (st, pd, oshape))
    num = idx[None, :, 1:] + pd_t - off[:, None, :]                                    # [K, N, 3]
    o = torch.div(num, st_t, rounding_mode="floor")
    ok = ((num % st_t == 0) & (num >= 0) & (o < os_t)).all(-1)
    b = idx[:, 0][None, :].expand_as(ok)
    D, H, W = oshape
    keys = ((b * D + o[..., 0]) * H + o[..., 1]) * W + o[..., 2]
    ukeys = torch.unique(keys[ok])                                                     # ascending linear index (spconv GPU path)
    out_idx = torch.stack([ukeys // (D * H * W), (ukeys // (H * W)) % D, (ukeys // W) % H, ukeys % W], 1)
    q = out_idx[None, :, 1:] * st_t - pd_t + off[:, None, :]
    return out_idx, oshape, _table(idx, spatial_shape, q, out_idx[:, 0])


def invert_rulebook(nbr_down, n_fine):
    K, M = nbr_down.shape
    up = torch.full((K, n_fine), -1, dtype=nbr_down.dtype, device=nbr_down.device)
    k, j = torch.nonzero(nbr_down >= 0, as_tuple=True)
    up[k, nbr_down[k, j]] = j
    return up


def sparse_conv(features, weight, nbr):
    """spconv 1.x indice_conv: per offset gather -> mm -> scatter-add (one index_select, one mm, one index_add_ each)."""
    K = nbr.shape[0]
    w = weight.reshape(K, weight.shape[-2], weight.shape[-1])
    out = torch.zeros(nbr.shape[1], w.shape[2], dtype=features.dtype, device=features.device)
    for k in range(K):
        j = torch.nonzero(nbr[k] >= 0).squeeze(1)
        if j.numel():
            out.index_add_(0, j, features.index_select(0, nbr[k, j]) @ w[k])
    return out


def three_nn(unknown, known, chunk=1 << 24):
    """3 nearest ``known`` per ``unknown`` row: fp32 (dx*dx + dy*dy) + dz*dz, ascending (distance, index).  The reference
    kernel (interpolate_gpu.cu:16-59) scans all M candidates per point in one thread; here a chunked distance matrix + topk
    (ties between equal distances may order differently from the sequential scan - a baseline, not the parity oracle)."""
    N, M = unknown.shape[0], known.shape[0]
    if unknown.is_cuda and M > 0 and N > 0:
        from . import ref_pointnet2 as rp
        if rp.available():            # the REAL reference kernel (three_nn_kernel_fast, brute force O(N*M), one thread per point)
            dist, idx = rp.three_nn(unknown[None, :, :3].contiguous(), known[None, :, :3].contiguous())
            return dist[0] * dist[0], idx[0].long()
    d2o = torch.full((N, 3), float("inf"), dtype=torch.float32, device=unknown.device)
    ido = torch.zeros((N, 3), dtype=torch.long, device=unknown.device)
    if M == 0 or N == 0:
        return d2o, ido
    kk = min(3, M)
    step = max(1, chunk // M)
    for s in range(0, N, step):
        u = unknown[s:s + step]
        dx = u[:, None, 0] - known[None, :, 0]
        dy = u[:, None, 1] - known[None, :, 1]
        dz = u[:, None, 2] - known[None, :, 2]
        d = (dx * dx + dy * dy) + dz * dz
        v, i = torch.topk(d, kk, dim=1, largest=False, sorted=True)
        d2o[s:s + step, :kk], ido[s:s + step, :kk] = v, i
    return d2o, ido
