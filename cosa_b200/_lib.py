"""ctypes binding of ``libcosa_b200.so`` (the C-ABI declared in ``include/cosa_b200.h``).

There is no CPU fallback and no alternative backend: if the shared library is missing, or a tensor is
not a CUDA tensor, the call fails loudly.  PyTorch is used only for device memory and streams.
"""
import ctypes
import os

import torch

from . import _lazy

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcosa_b200.so")

_c_int, _c_float, _c_size_t, _c_ll = ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_longlong
_vp = ctypes.c_void_p

# name -> (restype, argtypes); kept in step with include/cosa_b200.h (tests/test_abi.py checks the symbols)
_SIGNATURES = {
    "cosa_abi_version": (_c_int, []),
    "cosa_strerror": (ctypes.c_char_p, [_c_int]),
    "cosa_launch_count": (ctypes.c_ulonglong, []),
    "cosa_profile_begin": (None, []),
    "cosa_profile_end": (_c_int, [ctypes.c_char_p, _c_size_t]),
    "cosa_par_ws_bytes": (_c_size_t, [_c_int] * 5),
    "cosa_par_forward": (_c_int, [_vp, _vp, _vp] + [_c_int] * 6 + [_vp, _c_int, _c_int, _vp, _c_size_t, _vp]),
    "cosa_par_set_step_mode": (_c_int, [ctypes.c_char_p]),
    "cosa_par_affinity": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _vp, _c_int, _vp]),
    "cosa_cam_normalize": (_c_int, [_vp, _c_int, _vp, _c_int, _c_ll, _vp, _vp]),
    "cosa_multi_scale_cam_merge": (_c_int, [_vp, _vp, _vp, _c_int, _vp] + [_c_int] * 4 + [_vp, _vp]),
    "cosa_multi_scale_cam_merge_valid": (_c_int, [_vp, _vp, _vp, _c_int, _vp, _vp] + [_c_int] * 4 + [_vp, _vp]),
    "cosa_multi_scale_cam_merge_present": (_c_int, [_vp, _vp, _vp, _c_int, _vp, _vp] + [_c_int] * 4 + [_vp, _vp]),
    "cosa_multi_scale_seg_merge": (_c_int, [_vp, _vp, _vp, _c_int, _vp] + [_c_int] * 4 + [_vp]),
    "cosa_seg_loss_stats_bytes": (_c_size_t, []),
    "cosa_seg_loss_forward": (_c_int, [_vp, _vp, _c_float, _c_int, _vp, _vp] + [_c_int] * 4 + [_vp]),
    "cosa_seg_loss_backward": (_c_int, [_vp, _vp, _vp, _vp, _c_float, _c_int, _vp] + [_c_int] * 4 + [_vp]),
    "cosa_seg_refine_by_label": (_c_int, [_vp, _vp, _c_float, _c_int, _vp] + [_c_int] * 4 + [_vp]),
    "cosa_cam_loss_forward": (_c_int, [_vp, _vp, _c_int, _vp, _vp, _vp] + [_c_int] * 6 + [_vp]),
    "cosa_cam_loss_backward": (_c_int, [_vp, _vp, _vp, _c_int, _vp] + [_c_int] * 4 + [_vp]),
    "cosa_upsample_bilinear": (_c_int, [_vp, _vp, _c_ll, _c_int, _c_int, _c_int, _c_int, _vp]),
    "cosa_upsample_bilinear_backward_ws_bytes": (_c_size_t, [_c_ll, _c_int, _c_int]),
    "cosa_upsample_bilinear_backward": (_c_int, [_vp, _vp, _c_ll, _c_int, _c_int, _c_int, _c_int, _vp, _c_size_t, _vp]),
    "cosa_denormalize_img": (_c_int, [_vp, _vp, _c_int, _c_ll, _vp, _vp, _vp]),
    "cosa_cam_validation": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _c_ll, _vp]),
    "cosa_cam_to_label": (_c_int, [_vp] * 5 + [_c_int] * 4 + [_c_float] * 3 + [_c_int, _c_ll, _vp]),
    "cosa_cam2mask_ws_bytes": (_c_size_t, [_c_int] * 7),
    "cosa_cam2mask_ws_bytes_ex": (_c_size_t, [_c_int] * 8),
    "cosa_cam2mask": (_c_int, [_vp] * 4 + [_c_float] * 3 + [_c_int, _c_int, _vp, _c_int, _c_int] + [_vp] * 3 +
                      [_c_int] * 4 + [_vp, _c_size_t, _vp]),
    "cosa_cam2mask_flags": (_c_int, [_vp] * 4 + [_c_float] * 3 + [_c_int, _c_int, _vp, _c_int, _c_int] + [_vp] * 3 +
                            [_c_int] * 4 + [_vp, _c_size_t, _c_int, _vp]),
    "cosa_cam2mask_ex": (_c_int, [_vp] * 4 + [_c_float] * 3 + [_c_int, _c_int, _vp, _c_int, _c_int] + [_vp] * 3 +
                         [_c_int] * 4 + [_vp, _c_size_t, _c_int, _vp, _vp, _vp]),
    "cosa_upsample_argmax": (_c_int, [_vp, _vp, _vp] + [_c_int] * 6 + [_vp]),
    "cosa_bilateral_ws_bytes": (_c_size_t, [_c_int] * 4),
    "cosa_bilateralfilter_batch": (_c_int, [_vp, _vp, _vp] + [_c_int] * 4 + [_c_float] * 2 + [_vp, _c_size_t, _vp]),
    "cosa_bilateralfilter_batch_host": (_c_int, [_vp, _vp, _vp] + [_c_int] * 4 + [_c_float] * 2),
    "cosa_bilateral_stats": (_c_int, [_vp] + [_c_int] * 4 + [ctypes.POINTER(_c_ll), _vp]),
    "cosa_crf_inference_ws_bytes": (_c_size_t, [_c_int] * 4),
    "cosa_crf_inference": (_c_int, [_vp] * 3 + [_c_int] * 5 + [_c_float] * 5 + [_vp, _c_size_t, _vp]),
    "cosa_dense_energy_ws_bytes": (_c_size_t, [_c_int] * 4),
    "cosa_dense_energy_forward": (_c_int, [_vp] * 4 + [_c_float] * 2 + [_vp, _vp] + [_c_int] * 4 +
                                  [_vp, _c_size_t, _vp]),
    "cosa_dense_energy_backward": (_c_int, [_vp] * 4 + [_c_int] * 4 + [_vp]),
    "cosa_energy_loss_ws_bytes": (_c_size_t, [_c_int] * 4),
    "cosa_energy_loss_saved_bytes": (_c_size_t, [_c_int] * 4),
    "cosa_energy_loss_ws_bytes_ex": (_c_size_t, [_c_int] * 5),
    "cosa_energy_loss_lattice_offset": (_c_size_t, [_c_int] * 4),
    "cosa_energy_loss_prebuild_ex": (_c_int, [_vp] * 3 + [_c_float] * 2 + [_c_int] * 4 + [_vp, _c_size_t, _c_int, _vp]),
    "cosa_energy_loss_forward": (_c_int, [_vp] * 6 + [_c_float] * 3 + [_vp, _vp] + [_c_int] * 4 +
                                 [_vp, _c_size_t, _vp]),
    "cosa_energy_loss_prebuild": (_c_int, [_vp] * 3 + [_c_float] * 2 + [_c_int] * 4 + [_vp, _c_size_t, _vp]),
    "cosa_energy_loss_forward_flags": (_c_int, [_vp] * 6 + [_c_float] * 3 + [_vp, _vp] + [_c_int] * 4 +
                                       [_vp, _c_size_t, _c_int, _vp]),
    "cosa_energy_loss_forward_ev": (_c_int, [_vp] * 6 + [_c_float] * 3 + [_vp, _vp] + [_c_int] * 4 +
                                    [_vp, _c_size_t, _c_int, _vp, _vp]),
    "cosa_energy_loss_backward": (_c_int, [_vp] * 3 + [_c_float, _vp] + [_c_int] * 4 + [_vp]),
}

ABI_VERSION = 2
_lib = None


class CosaError(RuntimeError):
    pass


def load():
    """Load the CUDA library; raise if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CosaError(
                "cosa_b200: %s is missing - build it with `make -C cosa_b200/csrc` (needs nvcc). "
                "There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.cosa_abi_version() != ABI_VERSION:
            raise CosaError("cosa_b200: ABI version mismatch")
        _lib = lib
    return _lib


def check(code):
    if code != 0:
        raise CosaError("cosa_b200 kernel library error %d: %s" % (code, load().cosa_strerror(code).decode()))


def launch_count():
    return int(load().cosa_launch_count())


def profile_begin():
    load().cosa_profile_begin()


def profile_end():
    """{kernel name: (launches, total milliseconds)} for the launches since ``profile_begin``."""
    buf = ctypes.create_string_buffer(1 << 16)
    load().cosa_profile_end(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, count, ms = line.split()
        out[name] = (int(count), float(ms))
    return out


def stream_ptr():
    return _vp(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def dev_f32(t, what):
    """A contiguous float32 CUDA view of ``t``; refuses CPU tensors (no CPU path in this package)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % what)
    t = _lazy.plain(t)
    if not t.is_cuda:
        raise CosaError("cosa_b200: %s must be a CUDA tensor - this package has no CPU fallback" % what)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ---- stream-keyed scratch ---------------------------------------------------------------------------
_scratch = {}
_scratch_uses = {}     # per (device, stream): how many times the buffer has been handed out


def workspace(nbytes, device):
    """A cached uint8 scratch buffer of at least ``nbytes`` for the current stream of ``device``.

    Work on one stream is ordered, so successive calls may share the buffer; different streams get
    different buffers.  The buffer only grows (sized for the largest batch seen).
    """
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _scratch.pop(key, None)
        buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
        _scratch[key] = buf
    _scratch_uses[key] = _scratch_uses.get(key, 0) + 1
    return buf


def workspace_uses(device):
    """Number of times the current stream's scratch buffer has been handed out (every hand-out may overwrite it)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    return _scratch_uses.get(key, 0)


def release_workspaces():
    _scratch.clear()


class ResolvedBoxes:
    """Boxes already resolved for a given (B, H, W) and resident on the device (see ``resolve_boxes``): passing one
    as ``img_boxes`` / ``img_box`` skips the per-call host work and the small upload."""

    def __init__(self, tensor, B, H, W):
        self.tensor, self.B, self.H, self.W = tensor, B, H, W


def resolve_boxes(img_boxes, B, H, W, device):
    """[B,4] int32 device tensor of (y0, y1, x0, x1) after Python slice resolution.

    The reference indexes ``t[b, c0:c1, c2:c3]`` with whatever the loader gives it: an int16 tensor, or a
    list such as ``[[0, -1, 0, -1]]`` at eval time (negative ends drop the last row/column).  Images without
    a box keep ``ignore_index`` everywhere (``enumerate(img_boxes)`` simply stops).
    """
    if isinstance(img_boxes, ResolvedBoxes):
        if (img_boxes.B, img_boxes.H, img_boxes.W) != (B, H, W) or img_boxes.tensor.device != torch.device(device):
            raise CosaError("ResolvedBoxes were made for another batch geometry or device")
        return img_boxes.tensor
    if isinstance(img_boxes, torch.Tensor):
        rows = img_boxes.detach().cpu().tolist()
    else:
        rows = [[int(v) for v in coord] for coord in img_boxes]
    out = []
    for b in range(B):
        if b < len(rows):
            c = rows[b]
            y0, y1, _ = slice(int(c[0]), int(c[1])).indices(H)
            x0, x1, _ = slice(int(c[2]), int(c[3])).indices(W)
            out.append([y0, max(y1, y0), x0, max(x1, x0)])
        else:
            out.append([0, 0, 0, 0])
    t = torch.tensor(out, dtype=torch.int32)
    return t.to(device, non_blocking=True)
