"""ctypes bindings for the two CPU lattices (ORACLE — test infrastructure only).

* ``oracle_*``  -> oracle/liblattice_oracle.so, the plain-C restatement (lattice_oracle.c)
* ``ref_*``     -> oracle/_ref/libbf_ref.so, the UNMODIFIED reference C++
                   (utils/bilateralfilter/bilateralfilter.cpp:22-55) compiled by oracle/Makefile;
                   the SWIG glue (bilateralfilter.i) is replaced by calling the mangled C++ symbols.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liblattice_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libbf_ref.so")
_fp = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_ORACLE_SO) or (
            os.path.getmtime(_ORACLE_SO) < os.path.getmtime(os.path.join(_HERE, "lattice_oracle.c"))):
        subprocess.check_call(["make", "-C", _HERE, "liblattice_oracle.so"], stdout=subprocess.DEVNULL)
    if (force or not os.path.exists(_REF_SO)) and os.path.isdir("/root/reference/utils/bilateralfilter"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_oracle = None
_ref = None


def _oracle_lib():
    global _oracle
    if _oracle is None:
        build()
        lib = ctypes.CDLL(_ORACLE_SO)
        lib.cosa_oracle_bilateralfilter.restype = ctypes.c_int
        lib.cosa_oracle_bilateralfilter.argtypes = [_fp, _fp, _fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_float, ctypes.c_float]
        lib.cosa_oracle_bilateralfilter_batch.restype = None
        lib.cosa_oracle_bilateralfilter_batch.argtypes = [_fp, _fp, _fp] + [ctypes.c_int] * 4 + [ctypes.c_float] * 2
        lib.cosa_oracle_lattice_embed.restype = ctypes.c_int
        lib.cosa_oracle_lattice_embed.argtypes = [_fp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                                  ctypes.POINTER(ctypes.c_int), _fp,
                                                  ctypes.POINTER(ctypes.c_int16), ctypes.c_int]
        _oracle = lib
    return _oracle


def have_ref():
    return os.path.exists(_REF_SO)


def _ref_lib():
    global _ref
    if _ref is None:
        if not have_ref():
            build()
        lib = ctypes.CDLL(_REF_SO)
        # void bilateralfilter_batch(float*,int,float*,int,float*,int,int N,int K,int H,int W,float,float)
        f = lib._Z21bilateralfilter_batchPfiS_iS_iiiiiff
        f.restype = None
        f.argtypes = [_fp, ctypes.c_int, _fp, ctypes.c_int, _fp, ctypes.c_int] + [ctypes.c_int] * 4 + [ctypes.c_float] * 2
        _ref = lib
    return _ref


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))


def _ptr(a):
    return a.ctypes.data_as(_fp)


def oracle_bilateralfilter_batch(images, ins, outs, N, K, H, W, sigmargb, sigmaxy):
    """Same 9-argument call shape as the SWIG module (utils/seg_helper.py:887); writes ``outs`` in place."""
    images, ins = _f32(images), _f32(ins)
    assert outs.dtype == np.float32 and outs.flags.c_contiguous
    _oracle_lib().cosa_oracle_bilateralfilter_batch(_ptr(images), _ptr(ins), _ptr(outs), N, K, H, W,
                                                    float(sigmargb), float(sigmaxy))


def ref_bilateralfilter_batch(images, ins, outs, N, K, H, W, sigmargb, sigmaxy):
    images, ins = _f32(images), _f32(ins)
    assert outs.dtype == np.float32 and outs.flags.c_contiguous
    _ref_lib()._Z21bilateralfilter_batchPfiS_iS_iiiiiff(
        _ptr(images), images.size, _ptr(ins), ins.size, _ptr(outs), outs.size, N, K, H, W,
        float(sigmargb), float(sigmaxy))


def cpu_bilateralfilter_batch(*args):
    """The reference build when it is present (kind 'reference'), else the C port (kind 'port')."""
    return (ref_bilateralfilter_batch if have_ref() else oracle_bilateralfilter_batch)(*args)


def oracle_lattice_embed(image, H, W, sigmargb, sigmaxy):
    """Embedding of one planar RGB image: (offsets[n,6] int32, bary[n,6] f32, vkeys[M,5] int16)."""
    image = _f32(image)
    n = H * W
    cap = 6 * (n + 4)
    offsets = np.zeros((n, 6), np.int32)
    bary = np.zeros((n, 6), np.float32)
    vkeys = np.zeros((cap, 5), np.int16)
    M = _oracle_lib().cosa_oracle_lattice_embed(
        _ptr(image), H, W, float(sigmargb), float(sigmaxy),
        offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _ptr(bary),
        vkeys.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)), cap)
    assert M >= 0
    return offsets, bary, vkeys[:M].copy()
