"""Host side of the tcgen05 gather-GEMM (csrc/gather_gemm.cu): weight packing + launch helper."""
import os

import torch

from . import capi, ops as _ops


def round_tf32(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest (ties away) fp32 -> tf32, kept in an fp32 container (== cvt.rna.tf32.f32)."""
    xi = x.contiguous().view(torch.int32)
    r = (xi + 0x1000) & ~0x1FFF
    out = r.view(torch.float32)
    return torch.where(torch.isfinite(x), out, x)


def pad_to(v, m):
    return (v + m - 1) // m * m


# bench.py instrumentation: PROFILE = list -> CUDA events + static sizes per launch; COUNT = list -> rulebook pair counts
PROFILE = None
COUNT = None
DEBUG_SKIP = 0      # development only (see ls3d_gemm_args.debug_skip)


# Arithmetic of every gather-GEMM launch (csrc/gather_gemm*.cu): "bf16x3" - operands split into bf16 hi + lo on the fly,
# x_hi.W_hi + x_hi.W_lo + x_lo.W_hi with fp32 accumulate (~2^-17 relative error per product; activations stay exact fp32 in
# HBM).  ``ls3d_gemm_args.precise`` = 2 names it on the C ABI (the round-1 TF32 / 3xTF32 engines 0 / 1 are retired).
PRECISE = 2
# Sparse launches (nbr given) of the bf16x3 engine run on the gather-once kernel (csrc/gather_gemm_once.cu) with the tile
# plan of their rulebook; False = per-pair gather kernel (csrc/gather_gemm_bf16x3.cu), which also serves the dense Linears.
USE_PLAN = os.environ.get("LS3D_USE_PLAN", "1") == "1"
ENGINE_NAME = "gather_gemm_once_kernel" if USE_PLAN else "gather_gemm_bf16x3_kernel"      # bench.py's roofline label


def trunc_tf32(x: torch.Tensor) -> torch.Tensor:
    """What the tensor core does to an fp32 operand under kind::tf32: drop the 13 low mantissa bits."""
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


class PackedWeight:
    """W[koff][cin][cout] (spconv layout, reference scn_unet.py weight [kz,ky,kx,Cin,Cout]) as the bf16x3 weight image of
    include/ls3d.h (ls3d_gemm_pack_bf16x3): one swizzled [W_hi ; W_lo] block per (offset, 32-channel chunk)."""

    def __init__(self, w_kio: torch.Tensor):
        assert w_kio.dim() == 3
        koff, cin, cout = w_kio.shape
        self.koff, self.cin, self.cout = koff, cin, cout
        self.precise = 2
        self.cin_pad = pad_to(cin, 16)
        self.n_pad = pad_to(cout, 16)
        # CPU tensors (host-side tests of the layout) go through the tensor-op restatement of the same image
        buf = self._pack_bf16x3_device(w_kio) if w_kio.is_cuda else self._pack_bf16x3(w_kio.float().permute(0, 2, 1))
        self.data = buf.contiguous()

    def _pack_bf16x3_device(self, w_kio):
        """ls3d_gemm_pack_bf16x3 (the C-ABI packer, csrc/gather_gemm_bf16x3.cu): the layout documented in include/ls3d.h."""
        import ctypes
        koff, cin, cout = w_kio.shape
        nbytes, cp, npd = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int32()
        capi.check(capi.lib().ls3d_gemm_packed_bytes(koff, cin, cout, ctypes.byref(nbytes), ctypes.byref(cp), ctypes.byref(npd)),
                   "ls3d_gemm_packed_bytes")
        assert cp.value == self.cin_pad and npd.value == self.n_pad
        buf = torch.empty(nbytes.value // 2, dtype=torch.bfloat16, device=w_kio.device)
        w = w_kio.detach().float().contiguous()
        capi.check(capi.lib().ls3d_gemm_pack_bf16x3(capi.ptr(w), koff, cin, cout, capi.ptr(buf), capi.stream_ptr()),
                   "ls3d_gemm_pack_bf16x3")
        return buf

    def _pack_bf16x3(self, wt):
        """One contiguous block of n_pad * 128 bytes per (offset, 32-channel chunk), stored as the swizzled shared-memory
        image so that a single cp.async.bulk lands a ready tcgen05 K-major B tile (csrc/gather_gemm_bf16x3.cu):
          n_pad <= 96 ("stacked"): 2 n_pad rows of 64 bytes, rows [0, n) = bf16 hi of W[n, 32c..32c+31], rows [n, 2n) = bf16
                        lo; SWIZZLE_64B: 16-byte unit u of row r sits at unit u ^ ((r >> 1) & 3);
          n_pad  > 96 : n_pad rows of 128 bytes = [hi (32) | lo (32)]; SWIZZLE_128B: unit u of row r at u ^ (r & 7)."""
        koff, cout, cin = wt.shape
        nchunk = (self.cin_pad + 31) // 32
        n_pad = self.n_pad
        full = torch.zeros(koff, n_pad, nchunk * 32, dtype=torch.float32, device=wt.device)
        full[:, :cout, :cin] = wt
        hi = full.to(torch.bfloat16)
        lo = (full - hi.float()).to(torch.bfloat16)
        hi = hi.view(koff, n_pad, nchunk, 32).permute(0, 2, 1, 3)                          # [koff, nchunk, n, 32]
        lo = lo.view(koff, n_pad, nchunk, 32).permute(0, 2, 1, 3)
        if n_pad <= 96:
            rows = torch.cat([hi, lo], dim=2).reshape(koff, nchunk, 2 * n_pad, 4, 8)      # [.., row, 16B unit, 8 bf16]
            r = torch.arange(2 * n_pad, device=wt.device)
            u = torch.arange(4, device=wt.device)
            src = u[None, :] ^ ((r[:, None] >> 1) & 3)
        else:
            rows = torch.cat([hi, lo], dim=3).reshape(koff, nchunk, n_pad, 8, 8)
            r = torch.arange(n_pad, device=wt.device)
            u = torch.arange(8, device=wt.device)
            src = u[None, :] ^ (r[:, None] & 7)
        return rows[:, :, r[:, None], src].contiguous()

    @staticmethod
    def from_linear(weight_oi: torch.Tensor):
        """nn.Linear / Conv1d(k=1) weight [cout, cin] -> koff = 1."""
        return PackedWeight(weight_oi.t().unsqueeze(0))


def run(x0, pw: PackedWeight, *, x1=None, nbr=None, m_out=None, scale=None, shift=None, relu=False,
        res=None, res_mode=0, red=None, ln=(), ln_eps=1e-5, attn=None, row_mask=None, out=None, round_out=True):
    """Launch ls3d_gather_gemm.  ``x0``/``x1`` are [rows, C] fp32 row-major (row stride may exceed C).

    attn = dict(k=[F,H,L,24], v=[F,H,L,24], frame_off=int32[F], scale=float) selects the attention
    epilogue.  ``red`` = (t0, t1) enables the channel-reduction add.  ``ln`` = up to two (gamma, beta).
    """
    a = capi.GemmArgs()
    assert x0.dtype == torch.float32 and x0.stride(1) == 1
    c0 = x0.shape[1]
    c1 = 0
    a.in0, a.ld0, a.c0 = capi.ptr(x0), x0.stride(0), c0
    if x1 is not None:
        assert x1.dtype == torch.float32 and x1.stride(1) == 1
        c1 = x1.shape[1]
        a.in1, a.ld1 = capi.ptr(x1), x1.stride(0)
    a.c1 = c1
    assert c0 + c1 == pw.cin or (c0 + c1 <= pw.cin_pad and c0 + c1 >= pw.cin), (c0, c1, pw.cin)
    if nbr is not None:
        assert nbr.dtype == torch.int32 and nbr.is_contiguous() and nbr.shape[0] == pw.koff
        m = nbr.shape[1]
        a.nbr = capi.ptr(nbr)
        if USE_PLAN and pw.precise == 2 and attn is None and m > 0:
            plan = _ops.tile_plan(nbr)
            assert plan.m_out == m and plan.koff == pw.koff
            a.plan_hdr, a.plan_local, a.plan_pool = capi.ptr(plan.hdr), capi.ptr(plan.local), capi.ptr(plan.pool)
    else:
        m = x0.shape[0]
    if m_out is not None:
        m = m_out
    a.koff, a.m_out = pw.koff, m
    a.w, a.cin_pad, a.n_pad, a.cout = capi.ptr(pw.data), pw.cin_pad, pw.n_pad, pw.cout
    a.scale, a.shift, a.relu = capi.ptr(scale), capi.ptr(shift), int(relu)
    if res is not None:
        a.res, a.ld_res, a.res_mode = capi.ptr(res), res.stride(0), res_mode
    if red is not None:
        a.red0, a.red1 = capi.ptr(red[0]), capi.ptr(red[1])
        a.ld_red0, a.ld_red1, a.red_c = red[0].stride(0), red[1].stride(0), red[0].shape[1]
    a.n_ln = len(ln)
    if len(ln) > 0:
        a.ln_g0, a.ln_b0 = capi.ptr(ln[0][0]), capi.ptr(ln[0][1])
    if len(ln) > 1:
        a.ln_g1, a.ln_b1 = capi.ptr(ln[1][0]), capi.ptr(ln[1][1])
    a.ln_eps = ln_eps
    if attn is not None:
        a.epi = 1
        a.attn_k, a.attn_v = capi.ptr(attn["k"]), capi.ptr(attn["v"])
        a.frame_off = capi.ptr(attn["frame_off"])
        a.n_frames, a.n_head, a.n_tok = attn["k"].shape[0], attn["k"].shape[1], attn["k"].shape[2]
        a.attn_scale = attn["scale"]
    if row_mask is not None:
        a.row_mask, a.ld_mask = capi.ptr(row_mask), row_mask.stride(0)
    if out is None:
        out = torch.empty(m, pw.cout, dtype=torch.float32, device=x0.device)
    a.out, a.ld_out = capi.ptr(out), out.stride(0)
    a.round_out = 0            # tf32 rounding of stored activations belonged to the retired single-pass engine
    a.precise = int(pw.precise)
    a.debug_skip = DEBUG_SKIP
    if COUNT is not None:
        COUNT.append(int((nbr >= 0).sum()) if nbr is not None else m)
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        capi.gather_gemm(a)
        e1.record()
        PROFILE.append(dict(e0=e0, e1=e1, sparse=nbr is not None, rows_in=x0.shape[0], m_out=m, cin=c0 + c1, cout=pw.cout,
                            koff=pw.koff))
    else:
        capi.gather_gemm(a)
    return out
