"""ORACLE - test infrastructure only.  CPU restatement of the dense-CRF mean-field inference the reference runs at
evaluation time through pydensecrf (utils/seg_helper.py:961-996 ``DenseCRF.__call__`` / ``crf_inference_infv2``,
:905-922 ``crf_inference_inf``; caller evaluation_engine.py:205-211).

**Parity unpinned against pydensecrf.**  pydensecrf is a third-party dependency that the reference neither vendors nor
pins (README.md:104 installs lucasb-eyer/pydensecrf git master) and it is absent from this image, so there is no
reference run of this function to compare with.  What IS pinned:
  * the two permutohedral filters (2-D spatial, 5-D bilateral): the generic-dimension C restatement
    (oracle/lattice_oracle.c, built for D = 5 and D = 2) is compared with the REFERENCE's own ``Permutohedral`` class
    (utils/bilateralfilter/permutohedral.cpp - the same Kraehenbuehl code base pydensecrf wraps) through
    oracle/ref_lattice_shim.cpp, and the results are committed as tests/golden/crf_inference.npz;
  * the update rule, restated from the published algorithm and pydensecrf's sources as documented upstream:
      unary_from_softmax(sm, clip=1e-5):  U = -log(clip(sm, 1e-5, 1))                    pydensecrf/utils.py
      DenseCRF::inference(n):  Q = expAndNormalize(-U); n times: tmp = -U - sum_k pairwise_k(Q); Q = expAndNormalize(tmp)
      PairwisePotential::apply = PottsCompatibility(w) o DenseKernel:  -w * K(Q)          densecrf/src/pairwise.cpp
      DenseKernel (NORMALIZE_SYMMETRIC, the default of addPairwiseGaussian/Bilateral):
          norm = 1 / sqrt(filter(1) + 1e-20);  K(Q) = norm * filter(norm * Q)
      features: Gaussian (x / sxy, y / sxy); bilateral (x / sxy, y / sxy, R / srgb, G / srgb, B / srgb)   densecrf.cpp
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_fp = ctypes.POINTER(ctypes.c_float)
_libs = {}


def _lib(d):
    """liblattice_oracle.so (D = 5) / liblattice_oracle_d2.so (D = 2), built on demand."""
    if d not in _libs:
        name = "liblattice_oracle.so" if d == 5 else "liblattice_oracle_d%d.so" % d
        path = os.path.join(_HERE, name)
        src = os.path.join(_HERE, "lattice_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, name], stdout=subprocess.DEVNULL)
        lib = ctypes.CDLL(path)
        lib.cosa_oracle_filter_features.restype = ctypes.c_int
        lib.cosa_oracle_filter_features.argtypes = [_fp, ctypes.c_int, _fp, _fp, ctypes.c_int]
        assert lib.cosa_oracle_lattice_dim() == d
        _libs[d] = lib
    return _libs[d]


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_lattice.so"))


def ref_filter(features, values):
    """The REFERENCE's Permutohedral class (any d) through oracle/ref_lattice_shim.cpp; build container only."""
    lib = ctypes.CDLL(os.path.join(_HERE, "_ref", "libref_lattice.so"))
    lib.ref_lattice_filter.argtypes = [_fp, ctypes.c_int, ctypes.c_int, _fp, _fp, ctypes.c_int]
    feat = np.ascontiguousarray(features, dtype=np.float32)
    vals = np.ascontiguousarray(values, dtype=np.float32)
    n, d = feat.shape
    out = np.zeros_like(vals)
    lib.ref_lattice_filter(feat.ctypes.data_as(_fp), d, n, vals.ctypes.data_as(_fp), out.ctypes.data_as(_fp), vals.shape[0])
    return out


def oracle_filter(features, values):
    """Un-normalised permutohedral filter of ``values`` [K, n] over the lattice of ``features`` [n, d] (d = 2 or 5)."""
    feat = np.ascontiguousarray(features, dtype=np.float32)
    vals = np.ascontiguousarray(values, dtype=np.float32)
    n, d = feat.shape
    out = np.zeros_like(vals)
    _lib(d).cosa_oracle_filter_features(feat.ctypes.data_as(_fp), n, vals.ctypes.data_as(_fp), out.ctypes.data_as(_fp),
                                        vals.shape[0])
    return out


def gaussian_features(H, W, sxy):
    ys, xs = np.mgrid[0:H, 0:W]
    f = np.stack([xs.reshape(-1).astype(np.float32) / np.float32(sxy), ys.reshape(-1).astype(np.float32) / np.float32(sxy)], 1)
    return np.ascontiguousarray(f, dtype=np.float32)


def bilateral_features(image_hwc, sxy, srgb):
    H, W = image_hwc.shape[:2]
    ys, xs = np.mgrid[0:H, 0:W]
    rgb = image_hwc.reshape(-1, 3).astype(np.float32) / np.float32(srgb)
    f = np.concatenate([(xs.reshape(-1, 1).astype(np.float32) / np.float32(sxy)),
                        (ys.reshape(-1, 1).astype(np.float32) / np.float32(sxy)), rgb], 1)
    return np.ascontiguousarray(f, dtype=np.float32)


def _exp_and_normalize(t):
    t = t - t.max(axis=0, keepdims=True)
    e = np.exp(t, dtype=np.float32)
    return e / e.sum(axis=0, keepdims=True, dtype=np.float32)


def crf_inference(image_hwc, probs, iter_max, pos_w, pos_xy_std, bi_w, bi_xy_std, bi_rgb_std, filter_fn=oracle_filter):
    """Mean-field marginals [C, H, W] (float32) of the fully connected CRF; see the module docstring."""
    C, H, W = probs.shape
    n = H * W
    U = -np.log(np.clip(probs.reshape(C, n).astype(np.float32), 1e-5, 1.0)).astype(np.float32)
    Q = _exp_and_normalize(-U)
    if iter_max > 0:
        fg = gaussian_features(H, W, pos_xy_std)
        fb = bilateral_features(np.asarray(image_hwc), bi_xy_std, bi_rgb_std)
        ones = np.ones((1, n), dtype=np.float32)
        norm_g = (1.0 / np.sqrt(filter_fn(fg, ones) + np.float32(1e-20))).astype(np.float32)
        norm_b = (1.0 / np.sqrt(filter_fn(fb, ones) + np.float32(1e-20))).astype(np.float32)
        for _ in range(iter_max):
            kg = norm_g * filter_fn(fg, (Q * norm_g).astype(np.float32))
            kb = norm_b * filter_fn(fb, (Q * norm_b).astype(np.float32))
            Q = _exp_and_normalize((-U + np.float32(pos_w) * kg + np.float32(bi_w) * kb).astype(np.float32))
    return Q.reshape(C, H, W)
