"""Drop-in for the reference's SWIG module ``bilateralfilter`` (utils/bilateralfilter/bilateralfilter.i).

``bilateralfilter_batch(images, ins, outs, N, K, H, W, sigmargb, sigmaxy)`` keeps the 9-argument call of
utils/seg_helper.py:887: ``images``/``ins`` are 1-D array-likes converted to contiguous float32
(SWIG IN_ARRAY1), ``outs`` must already be a contiguous float32 ndarray and is written in place
(INPLACE_ARRAY1, else TypeError).  The filter itself runs on the GPU lattice (``cosa_bilateralfilter_batch_host``).
CUDA tensors are accepted too and then nothing leaves the device.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def bilateralfilter_batch(images, ins, outs, N, K, H, W, sigmargb, sigmaxy):
    lib = _lib.load()
    N, K, H, W = int(N), int(K), int(H), int(W)
    if isinstance(outs, torch.Tensor):
        return _device_call(lib, images, ins, outs, N, K, H, W, float(sigmargb), float(sigmaxy))
    if not (isinstance(outs, np.ndarray) and outs.dtype == np.float32 and outs.flags.c_contiguous
            and outs.flags.writeable):
        raise TypeError("outs must be a contiguous, writeable float32 numpy array (INPLACE_ARRAY1)")
    images = np.ascontiguousarray(np.asarray(images, dtype=np.float32).reshape(-1))
    ins = np.ascontiguousarray(np.asarray(ins, dtype=np.float32).reshape(-1))
    # the reference ignores the lengths (bilateralfilter.cpp:42-55); reading past a short buffer is refused here
    if images.size < N * 3 * H * W or ins.size < N * K * H * W or outs.size < N * K * H * W:
        raise ValueError("buffer shorter than N*K*H*W")
    fp = ctypes.c_void_p
    _lib.check(lib.cosa_bilateralfilter_batch_host(fp(images.ctypes.data), fp(ins.ctypes.data), fp(outs.ctypes.data),
                                                   N, K, H, W, float(sigmargb), float(sigmaxy)))


def bilateralfilter(image, in_, out, H, W, sigmargb, sigmaxy):
    """Single-image form (bilateralfilter.hpp:11); K is inferred from len(in_) like bilateralfilter.cpp:27."""
    K = int(np.asarray(in_).size // (H * W)) if not isinstance(in_, torch.Tensor) else int(in_.numel() // (H * W))
    return bilateralfilter_batch(image, in_, out, 1, K, H, W, sigmargb, sigmaxy)


def _device_call(lib, images, ins, outs, N, K, H, W, sigmargb, sigmaxy):
    images = _lib.dev_f32(images, "images")
    ins = _lib.dev_f32(ins, "ins")
    if not (outs.is_cuda and outs.dtype == torch.float32 and outs.is_contiguous()):
        raise TypeError("outs must be a contiguous float32 CUDA tensor")
    with torch.cuda.device(outs.device):
        nbytes = lib.cosa_bilateral_ws_bytes(N, K, H, W)
        ws = _lib.workspace(nbytes, outs.device)
        _lib.check(lib.cosa_bilateralfilter_batch(_lib.ptr(images), _lib.ptr(ins), _lib.ptr(outs), N, K, H, W,
                                                  sigmargb, sigmaxy, _lib.ptr(ws), nbytes, _lib.stream_ptr()))


def lattice_stats(N, K, H, W, device=None):
    """(M, key_range_error, table_capacity, max_probe) of the last device-side filter call with these shapes."""
    lib = _lib.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    with torch.cuda.device(device):
        nbytes = lib.cosa_bilateral_ws_bytes(N, K, H, W)
        ws = _lib.workspace(nbytes, device)
        stats = (ctypes.c_longlong * 4)()
        rc = lib.cosa_bilateral_stats(_lib.ptr(ws), N, K, H, W, stats, _lib.stream_ptr())
    if rc not in (0, -3):
        _lib.check(rc)
    return tuple(int(v) for v in stats)
