"""cosa_b200 - B200-native CAM -> pseudo-label refinement path of CoSA.

Drop-in replacements (same names, arguments and return conventions as the reference) whose work is done
by hand-written sm_100a CUDA kernels behind the C-ABI in ``include/cosa_b200.h``:

  ``PAR``                                           models/PAR.py:26-91
  ``cam_validation, cam_to_label, cam2mask, _refine_cams``   utils/seg_helper.py:515-551, 721-797
  ``cam_normalize``                                 utils/seg_helper.py:264-270
  ``denormalize_img``                               utils/torch_helper.py:354-367 (main.py:117)
  ``upsample_bilinear``                             main.py:167 (F.interpolate of the logits, with its adjoint)
  ``get_energy_loss, DenseEnergyLoss, DenseEnergyLossFunction``   utils/seg_helper.py:191-230, 864-903
  ``multi_scale_camseg`` (merge fused), ``seg_loss``, ``seg_refine_by_label``, ``cam_loss``   :232-275, 800-813, 553-602
  ``bilateralfilter.bilateralfilter_batch``         utils/bilateralfilter (SWIG module)
  ``DenseCRF, crf_inference_infv2, crf_inference_inf``   utils/seg_helper.py:905-922, 961-996 (dense-CRF inference)

Everything requires CUDA tensors; there is no CPU fallback and no second backend.
"""
from . import _lib  # noqa: F401
from .host_pipeline import GraphedStep, HostPipeline
from .par import PAR, get_kernel
from .seg_helper import (DenseCRF, crf_inference_batch, crf_inference_inf, crf_inference_infv2, DenseEnergyLoss, DenseEnergyLossFunction, _refine_cams, cam2mask, cam_normalize,
                         cam_to_label, cam_validation, denormalize_img, get_energy_loss, multi_scale_cam_merge, multi_scale_camseg,
                         multi_scale_seg_merge, cam_loss, seg_loss, seg_refine_by_label, upsample_bilinear)

__all__ = ["PAR", "get_kernel", "cam_validation", "cam_to_label", "cam2mask", "_refine_cams", "cam_normalize",
           "get_energy_loss", "DenseEnergyLoss", "DenseEnergyLossFunction", "multi_scale_camseg", "multi_scale_cam_merge",
           "multi_scale_seg_merge", "HostPipeline", "GraphedStep", "seg_loss",
           "seg_refine_by_label", "cam_loss", "denormalize_img", "upsample_bilinear", "DenseCRF", "crf_inference_infv2",
           "crf_inference_inf", "crf_inference_batch"]
