"""ORACLE — test infrastructure only.  Imports the REFERENCE's own Python hot path, unmodified, when /root/reference is
mounted (the build container; never the GPU box).

  * models/PAR.py is loaded by file path (importing ``models`` would pull in timm, absent here);
  * utils/seg_helper.py is imported with a ctypes shim module named ``bilateralfilter`` in front of the UNMODIFIED
    reference C++ (oracle/_ref/libbf_ref.so, built by oracle/Makefile) in place of the SWIG glue, and with
    ``pydensecrf`` stubbed (not on this path);
  * ``Tensor.cuda`` is neutralised when there is no GPU (seg_helper.py:230,880,901).
Nothing else is patched.  Used by tests/golden/make_golden.py (golden vectors) and by ``bench.py --impl reference``
(kind "reference": the reference's own code timed on the host cores, once, beside the port).
"""
import importlib.util
import os
import sys
import types

import torch

from . import lattice as olat

REF = "/root/reference"


def available():
    return os.path.isfile(os.path.join(REF, "utils", "seg_helper.py")) and os.path.isfile(
        os.path.join(REF, "models", "PAR.py"))


def load_reference():
    """(PAR module, seg_helper module, torch_helper module) of the reference, or None when it is not mounted."""
    if not available():
        return None
    olat.build()
    shim = types.ModuleType("bilateralfilter")
    shim.bilateralfilter_batch = olat.ref_bilateralfilter_batch
    shim.bilateralfilter = None
    sys.modules["bilateralfilter"] = shim
    for name in ("pydensecrf", "pydensecrf.densecrf", "pydensecrf.utils"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["pydensecrf.utils"].unary_from_softmax = None
    sys.modules["pydensecrf"].densecrf = sys.modules["pydensecrf.densecrf"]
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self

    def by_path(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    # utils/torch_helper.py (denormalize_img) imports its sibling ``misc`` and texttable (absent): both unused here
    pkg = types.ModuleType("utils")
    pkg.__path__ = []
    sys.modules.setdefault("utils", pkg)
    sys.modules.setdefault("utils.misc", types.ModuleType("utils.misc"))
    tt = types.ModuleType("texttable")
    tt.Texttable = None
    sys.modules.setdefault("texttable", tt)
    th = by_path("utils.torch_helper", os.path.join(REF, "utils", "torch_helper.py"))
    par = by_path("ref_PAR", os.path.join(REF, "models", "PAR.py"))
    sh = by_path("ref_seg_helper", os.path.join(REF, "utils", "seg_helper.py"))
    return par, sh, th
