"""PAR - pixel-adaptive refinement, same class/ctor/forward as the reference ``models/PAR.py:26-91``.

The whole forward (bilinear ``align_corners=True`` resize of the masks when needed, the 8*len(dilations)
neighbour affinity from RGB + position, and ``num_iter`` propagation steps) runs in the sm_100a kernels of
``csrc/par_kernels.cu`` through ``cosa_par_forward``.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib


_STEP_MODE = "chain"


def set_step_mode(name):
    """Select the propagation kernel: "chain" (default: all steps in one launch, tile-level step counters;
    "chain<G>" sets the images per group), "tile" (one launch per step) or "smem" (the generic per-step kernel);
    process-wide, for A/B runs and tests - see cosa_par_set_step_mode."""
    global _STEP_MODE
    _lib.check(_lib.load().cosa_par_set_step_mode(name.encode()))
    _STEP_MODE = name


def step_mode():
    """The propagation kernel selected by set_step_mode."""
    return _STEP_MODE


def lattice_prebuild_before_cam2mask():
    """Where a step should start the CRF lattice build on the second stream (DenseEnergyLoss.prebuild_lattice): True =
    before cam2mask, False = between cam2mask and get_energy_loss.  Measured on B200 (profiles/README.md): next to the
    per-step launches of the "tile" kernel the build fills the wave tails of PAR; next to the single long grid of the
    "chain" kernel it makes the step slower (2.10 -> 2.21 ms), so there it is started behind cam2mask and runs beside
    the HBM-bound softmax / gate kernel of get_energy_loss (the hand-over is inside cosa_energy_loss_forward_ev)."""
    return _STEP_MODE.startswith("tile")


def get_kernel():
    """The 8 one-hot 3x3 taps of the reference (models/PAR.py:10-24); kept as a buffer for state_dict parity."""
    weight = torch.zeros(8, 1, 3, 3)
    for i, (r, c) in enumerate(((0, 0), (0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1), (2, 2))):
        weight[i, 0, r, c] = 1
    return weight


class PAR(nn.Module):

    def __init__(self, dilations, num_iter):
        super().__init__()
        self.dilations = list(dilations)
        self.num_iter = num_iter
        self.register_buffer('kernel', get_kernel())
        self.pos = self.get_pos()
        self.dim = 2
        self.w1 = 0.3
        self.w2 = 0.01
        self._dil = (ctypes.c_int * len(self.dilations))(*[int(d) for d in self.dilations])
        self._shared = None          # state of shared_affinity(): None outside the context

    def shared_affinity(self):
        """Context manager for consecutive ``cam2mask(..., refine_model=self)`` calls on the SAME images (main.py:158
        and :191 label the CAMs and the auxiliary CAMs of one batch): the affinity of the first call is reused by the
        following ones instead of being recomputed.  The caller vouches for the images being the same; geometry,
        dilations and an untouched workspace are checked, and anything else falls back to recomputing."""
        par = self

        class _Ctx:
            def __enter__(self):
                par._shared = {"key": None, "uses": None}
                return par

            def __exit__(self, *exc):
                par._shared = None
                return False

        return _Ctx()

    def get_pos(self):
        """[1,1,8*len(dilations),1,1] neighbour distances (models/PAR.py:51-62); the kernels use the same table."""
        ker = torch.ones(1, 1, 8, 1, 1)
        for m in (0, 2, 5, 7):
            ker[0, 0, m, 0, 0] = np.sqrt(2)
        return torch.cat([ker * d for d in self.dilations], dim=2)

    def affinity(self, imgs):
        """The affinity tensor alone, [b, 8*len(dilations), h, w] (models/PAR.py:69-85)."""
        lib = _lib.load()
        imgs = _lib.dev_f32(imgs, "imgs")
        b, c, h, w = imgs.shape
        assert c == 3, "PAR expects RGB images"
        out = torch.empty((b, 8 * len(self.dilations), h, w), dtype=torch.float32, device=imgs.device)
        with torch.cuda.device(imgs.device):
            _lib.check(lib.cosa_par_affinity(_lib.ptr(imgs), _lib.ptr(out), b, h, w, self._dil, len(self.dilations),
                                             _lib.stream_ptr()))
        return out

    def forward(self, imgs, masks):
        lib = _lib.load()
        imgs = _lib.dev_f32(imgs, "imgs")
        masks = _lib.dev_f32(masks, "masks")
        b, c, h, w = imgs.shape
        assert c == 3, "PAR expects RGB images"
        bm, cm, hm, wm = masks.shape
        assert bm == b, "imgs and masks must share the batch size"
        out = torch.empty((b, cm, h, w), dtype=torch.float32, device=imgs.device)
        n_dil = len(self.dilations)
        with torch.cuda.device(imgs.device):
            nbytes = lib.cosa_par_ws_bytes(b, cm, h, w, n_dil)
            ws = _lib.workspace(nbytes, imgs.device)
            _lib.check(lib.cosa_par_forward(_lib.ptr(imgs), _lib.ptr(masks), _lib.ptr(out), b, cm, h, w, hm, wm,
                                            self._dil, n_dil, int(self.num_iter), _lib.ptr(ws), nbytes,
                                            _lib.stream_ptr()))
        return out
