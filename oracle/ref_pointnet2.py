"""ctypes binding of the REAL reference pointnet2 interpolation kernels (oracle/_ref/libpointnet2_ref.so, built by
oracle/Makefile from /root/reference/det3d/ops/pointnet2_batch/src/interpolate_gpu.cu, unmodified).

TEST INFRASTRUCTURE + GPU BASELINE ONLY.  Signatures follow the reference launchers (interpolate_gpu.h:12-33), which are
what pointnet2_utils.ThreeNN / ThreeInterpolate call (pointnet2_utils.py:76-153); like there, tensors are [B, N, 3] /
[B, C, M] contiguous and the kernels run on the legacy default stream (interpolate_gpu.cu:74,117)."""
import ctypes
import os

import torch

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libpointnet2_ref.so")
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(LIB)
        P, I = ctypes.c_void_p, ctypes.c_int
        _lib._Z29three_nn_kernel_launcher_fastiiiPKfS0_PfPi.argtypes = [I, I, I, P, P, P, P]
        _lib._Z38three_interpolate_kernel_launcher_fastiiiiPKfPKiS0_Pf.argtypes = [I, I, I, I, P, P, P, P]
        _lib._Z43three_interpolate_grad_kernel_launcher_fastiiiiPKfPKiS0_Pf.argtypes = [I, I, I, I, P, P, P, P]
    return _lib


def three_nn(unknown, known):
    """unknown [B, N, 3], known [B, M, 3] fp32 cuda -> (dist [B, N, 3] = sqrt(d2), idx [B, N, 3] int32), as ThreeNN.forward
    (pointnet2_utils.py:78-98) returns them."""
    assert unknown.is_cuda and unknown.is_contiguous() and known.is_contiguous()
    B, N, _ = unknown.shape
    m = known.shape[1]
    d2 = torch.empty(B, N, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(B, N, 3, dtype=torch.int32, device=unknown.device)
    torch.cuda.current_stream().synchronize()
    lib()._Z29three_nn_kernel_launcher_fastiiiPKfS0_PfPi(B, N, m, unknown.data_ptr(), known.data_ptr(), d2.data_ptr(),
                                                          idx.data_ptr())
    torch.cuda.synchronize()
    return torch.sqrt(d2), idx


def three_interpolate(features, idx, weight):
    """features [B, C, M], idx [B, N, 3] int32, weight [B, N, 3] -> [B, C, N] (ThreeInterpolate.forward, :111-132)."""
    features, idx, weight = features.contiguous(), idx.contiguous(), weight.contiguous()
    B, C, M = features.shape
    N = idx.shape[1]
    out = torch.empty(B, C, N, dtype=torch.float32, device=features.device)
    torch.cuda.current_stream().synchronize()
    lib()._Z38three_interpolate_kernel_launcher_fastiiiiPKfPKiS0_Pf(B, C, M, N, features.data_ptr(), idx.data_ptr(),
                                                                     weight.data_ptr(), out.data_ptr())
    torch.cuda.synchronize()
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """grad_out [B, C, N] -> grad_features [B, C, M] (ThreeInterpolate.backward, :134-153)."""
    grad_out, idx, weight = grad_out.contiguous(), idx.contiguous(), weight.contiguous()
    B, C, N = grad_out.shape
    g = torch.zeros(B, C, m, dtype=torch.float32, device=grad_out.device)
    torch.cuda.current_stream().synchronize()
    lib()._Z43three_interpolate_grad_kernel_launcher_fastiiiiPKfPKiS0_Pf(B, C, N, m, grad_out.data_ptr(), idx.data_ptr(),
                                                                         weight.data_ptr(), g.data_ptr())
    torch.cuda.synchronize()
    return g
