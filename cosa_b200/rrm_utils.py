"""Second import path of the dense-CRF loss classes, mirroring the reference's duplicate in
``utils/rrm_utils.py:352-416`` (same maths as ``utils/seg_helper.py``; the layer there calls
``F.interpolate`` without ``recompute_scale_factor``).  The rest of rrm_utils (numpy/pydensecrf label code,
``:23-79``) is dead code in the reference and out of scope.
"""
from . import seg_helper as _sh
from .seg_helper import DenseEnergyLossFunction  # noqa: F401


class DenseEnergyLoss(_sh.DenseEnergyLoss):
    recompute_scale_factor = False


_sh._FUSABLE_LAYERS.add(DenseEnergyLoss)
