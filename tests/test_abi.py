"""CPU: the C-ABI library loads, exports every symbol include/cosa_b200.h declares, and the host-side
logic (workspace sizing, box resolution, module surface) behaves.  No kernel is launched here."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "cosa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cosa_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from cosa_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from cosa_b200 import _lib
    names = header_symbols()
    assert len(names) >= 25
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libcosa_b200.so does not export " + n
    assert set(_lib._SIGNATURES) == set(names), set(_lib._SIGNATURES) ^ set(names)
    assert lib.cosa_abi_version() == _lib.ABI_VERSION == 2
    assert b"workspace" in lib.cosa_strerror(-2)


def test_library_is_sm100a_only():
    from cosa_b200 import _lib
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_workspace_sizes_are_monotone(lib):
    a = lib.cosa_bilateral_ws_bytes(4, 21, 224, 224)
    b = lib.cosa_bilateral_ws_bytes(32, 21, 224, 224)
    c = lib.cosa_bilateral_ws_bytes(32, 81, 224, 224)
    assert 0 < a < b < c < 16 << 30
    assert lib.cosa_bilateral_ws_bytes(100, 21, 224, 224) <= lib.cosa_bilateral_ws_bytes(64, 21, 224, 224)  # two chunks of 50
    assert lib.cosa_cam2mask_ws_bytes(32, 20, 448, 448, 2, 1, 6) > lib.cosa_cam2mask_ws_bytes(32, 20, 448, 448, 2, 0, 0)
    assert lib.cosa_energy_loss_ws_bytes(32, 21, 448, 448) > lib.cosa_dense_energy_ws_bytes(32, 21, 224, 224)
    assert lib.cosa_par_ws_bytes(1, 20, 224, 224, 6) > 48 * 224 * 224 * 4


def test_argument_errors_do_not_launch(lib):
    before = lib.cosa_launch_count()
    assert lib.cosa_cam_validation(None, None, None, 1, 1, 1, None) == -1
    assert lib.cosa_bilateralfilter_batch(None, None, None, 1, 1, 8, 8, 15.0, 50.0, None, 0, None) == -1
    assert lib.cosa_denormalize_img(None, None, 1, 64, None, None, None) == -1
    assert lib.cosa_upsample_bilinear(None, None, 1, 2, 2, 4, 4, None) == -1
    assert lib.cosa_upsample_bilinear_backward(None, None, 1, 2, 2, 4, 4, None, 0, None) == -1
    assert lib.cosa_multi_scale_cam_merge_valid(None, None, None, 1, None, None, 1, 1, 4, 4, None, None) == -1
    assert lib.cosa_cam2mask_flags(None, None, None, None, 0.7, 0.25, 255.0, 2, 0, None, 0, 0, None, None, None,
                                   1, 1, 4, 4, None, 0, 0, None) == -1
    assert lib.cosa_energy_loss_prebuild(None, None, None, 15.0, 50.0, 1, 21, 8, 8, None, 0, None) == -1
    assert lib.cosa_energy_loss_forward_flags(None, None, None, None, None, None, 1e-7, 15.0, 50.0, None, None, 1, 21,
                                              8, 8, None, 0, 2, None) == -1       # unknown flag bit
    assert lib.cosa_upsample_bilinear_backward_ws_bytes(42, 28, 448) == 42 * 28 * 448 * 2 * 4
    assert lib.cosa_launch_count() == before


def test_box_resolution_follows_python_slices():
    from cosa_b200._lib import resolve_boxes
    cpu = torch.device("cpu")
    got = resolve_boxes(torch.tensor([[0, 448, 0, 448], [16, 432, 32, 448]], dtype=torch.int16), 2, 448, 448, cpu)
    assert got.dtype == torch.int32 and got.tolist() == [[0, 448, 0, 448], [16, 432, 32, 448]]
    assert resolve_boxes([[0, -1, 0, -1]], 1, 100, 60, cpu).tolist() == [[0, 99, 0, 59]]      # evaluation_engine.py:134
    assert resolve_boxes([[10, 5, 0, 999]], 1, 100, 60, cpu).tolist() == [[10, 10, 0, 60]]    # empty rows, clamped cols
    assert resolve_boxes([[0, 4, 0, 4]], 3, 8, 8, cpu).tolist() == [[0, 4, 0, 4], [0, 0, 0, 0], [0, 0, 0, 0]]


def test_module_surface_matches_reference_names():
    import inspect
    import cosa_b200
    from cosa_b200 import bilateralfilter, rrm_utils, seg_helper
    assert list(inspect.signature(seg_helper.cam2mask).parameters)[:9] == [
        "images", "img_boxes", "cams", "cls_labels", "threshold_high", "threshold_low", "refine_model",
        "ignore_index", "downscale"]
    assert list(inspect.signature(seg_helper.cam_to_label).parameters) == [
        "cam", "cls_label", "img_box", "bkg_thre", "high_thre", "low_thre", "ignore_mid", "ignore_index"]
    assert list(inspect.signature(seg_helper.get_energy_loss).parameters) == [
        "img", "logit", "label", "img_box", "loss_layer", "mean", "std"]
    assert list(inspect.signature(seg_helper._refine_cams).parameters) == [
        "refine_model", "images", "cams", "valid_key", "orig_size"]
    assert list(inspect.signature(bilateralfilter.bilateralfilter_batch).parameters) == [
        "images", "ins", "outs", "N", "K", "H", "W", "sigmargb", "sigmaxy"]
    assert list(inspect.signature(cosa_b200.PAR.__init__).parameters) == ["self", "dilations", "num_iter"]
    layer = rrm_utils.DenseEnergyLoss(1e-7, 15, 100, 0.5)
    assert "sigma_rgb=15" in layer.extra_repr() and issubclass(rrm_utils.DenseEnergyLoss, seg_helper.DenseEnergyLoss)
    par = cosa_b200.PAR(dilations=[1, 2, 4, 8, 12, 24], num_iter=10)
    assert list(par.state_dict()) == ["kernel"] and par.pos.shape == (1, 1, 48, 1, 1)


def test_loss_layer_copies_leave_the_prebuild_scratch_behind():
    """DenseEnergyLoss.prebuild_lattice keeps a stream, a workspace and a pending event on the layer; copies and
    pickles of the layer (cosa_b200.GraphedStep takes a private copy) must not share or serialise them."""
    import copy
    import pickle
    import cosa_b200
    layer = cosa_b200.DenseEnergyLoss(1e-7, 15, 100, 0.5)
    layer.__dict__["_pre_state"] = {"stream": object()}
    layer.__dict__["_prebuilt"] = {"img_ptr": 1}
    for other in (copy.copy(layer), copy.deepcopy(layer), pickle.loads(pickle.dumps(layer))):
        assert type(other) is type(layer) and other.extra_repr() == layer.extra_repr()
        assert "_pre_state" not in other.__dict__ and "_prebuilt" not in other.__dict__
    assert "_pre_state" in layer.__dict__


def test_cpu_tensors_are_refused_not_emulated():
    import cosa_b200
    from cosa_b200._lib import CosaError
    with pytest.raises(CosaError):
        cosa_b200.cam_validation(torch.rand(1, 2, 4, 4), torch.ones(1, 2))
    with pytest.raises(CosaError):
        cosa_b200.get_energy_loss(torch.rand(1, 3, 8, 8), torch.rand(1, 4, 8, 8), torch.zeros(1, 8, 8), [[0, 8, 0, 8]],
                                  cosa_b200.DenseEnergyLoss(1e-7, 15, 100, 0.5))
    # the optional lattice prebuild declines CPU tensors (nothing to overlap with); get_energy_loss then refuses them
    assert cosa_b200.DenseEnergyLoss(1e-7, 15, 100, 0.5).prebuild_lattice(torch.rand(1, 3, 8, 8), 4) is False


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cosa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "synthetic.py", os.path.join(dirpath, f)


def test_synthetic_batch_is_seeded_and_shaped():
    from cosa_b200 import synthetic
    a = synthetic.synthetic_batch(B=2, C=21, H=64, W=64, n_fg=2, seed=3)
    b = synthetic.synthetic_batch(B=2, C=21, H=64, W=64, n_fg=2, seed=3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert a["cams"].shape == (2, 20, 64, 64) and a["logits"].shape == (2, 21, 64, 64)
    assert a["cls_label"].sum(1).tolist() == [2.0, 2.0]
    assert float(a["img_denorm"].max()) <= 1.0 and a["img_box"].dtype == torch.int16


def test_par_step_mode_switch_and_prebuild_policy():
    """cosa_par_set_step_mode is host-only: the names, the error for an unknown one, and the policy that follows the
    mode (where a step starts the second-stream lattice build)."""
    import cosa_b200
    from cosa_b200 import par as par_mod
    assert par_mod.step_mode() == "chain" and par_mod.lattice_prebuild_before_cam2mask() is False
    try:
        for name in ("tile", "smem", "chain16", "chain"):
            par_mod.set_step_mode(name)
            assert par_mod.step_mode() == name
            assert par_mod.lattice_prebuild_before_cam2mask() is (name == "tile")
        with pytest.raises(cosa_b200._lib.CosaError):
            par_mod.set_step_mode("nonsense")
        assert par_mod.step_mode() == "chain"          # a refused name changes nothing
    finally:
        par_mod.set_step_mode("chain")
