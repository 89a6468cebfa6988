"""CPU: the oracle (oracle/) against the golden vectors produced by the reference's own code.

Integer outputs (labels) must be identical.  Float outputs must be bit-identical when this host runs
the same torch build and CPU capability that made the fixtures (tests/golden/META.txt); on a different
host torch's vectorised exp/reductions may differ in the last ulps, so the check relaxes to 1e-6
relative there (the product tolerance is 1e-4).
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden, rel_inf, t
from oracle import lattice as olat
from oracle import reference_port as port


def _same_host_build():
    meta = dict(l.strip().split(": ", 1) for l in open(os.path.join(GOLDEN, "META.txt")))
    return meta["torch"] == torch.__version__ and meta["cpu_capability"] == torch.backends.cpu.get_cpu_capability()


def check_float(got, want, what):
    got = got.detach().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    if _same_host_build():
        assert np.array_equal(got, want), "%s: not bit-identical, rel=%g" % (what, rel_inf(got, want))
    else:
        assert rel_inf(got, want) <= 1e-6, what


def check_labels(got, want, what):
    got = got.detach().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    assert got.shape == want.shape, what
    bad = int((got != want).sum())
    if _same_host_build():
        assert bad == 0, "%s: %d label mismatches" % (what, bad)
    else:   # near-tie flips only
        assert bad <= 1e-4 * want.size, "%s: %d label mismatches" % (what, bad)


def test_par_matches_reference():
    g = load_golden("par")
    imgs, masks = t(g["imgs"]), t(g["masks"])
    check_float(port.par_forward(imgs, masks), g["out"], "PAR 10 iter")
    check_float(port.par_forward(imgs, masks, (1, 2, 4), 3), g["out_iter3_dil124"], "PAR 3 iter")
    check_float(port.par_forward(imgs, t(g["masks_lr"])), g["out_lr"], "PAR resized masks")
    assert rel_inf(port.par_position_affinity(port.DEFAULT_DILATIONS)[0, 0, :, 0, 0], g["pos_softmax"]) < 1e-6
    aff = port.par_affinity(imgs)
    assert abs(float(aff.sum(2).mean()) - 1.01) < 1e-5          # sum of affinities is 1 + w2


def test_normalize_and_validation():
    g = load_golden("normalize")
    check_float(port.normalize_cam([t(g["s0"]), t(g["s1"]), t(g["s2"])]), g["out"], "normalize")
    g = load_golden("cam_to_label")
    check_float(port.cam_validation(t(g["cam"]), t(g["cls_label"])), g["valid"], "cam_validation")


def test_denormalize_img():
    g = load_golden("denormalize")                      # every uint8 level in every channel + inexact inputs
    out = port.denormalize_img(t(g["simg"]))
    assert torch.equal(out, t(g["out"]))


def test_multi_scale_merge():
    g = load_golden("multi_scale")
    raw = lambda k: [t(g["raw_%s%d" % (k, i)]) for i in range(3)]
    cam, aux, seg = port.multi_scale_merge(raw("cam"), t(g["raw_aux2"]), raw("seg"), g["imgs"].shape[2:])
    check_float(cam, g["cam"], "merged cam")
    check_float(aux, g["cam_aux"], "merged aux cam (last scale only, seg_helper.py:258)")
    check_float(seg, g["seg"], "merged seg")


def test_losses_next_rows():
    """seg_loss, seg_refine_by_label, cam_loss against the reference's outputs."""
    g = load_golden("losses")
    sp = t(g["logits"]).requires_grad_(True)
    loss = port.seg_loss(sp, t(g["label"]), fg_alpha=0.5)
    loss.backward()
    check_float(loss, g["seg_loss"], "seg_loss")
    check_float(sp.grad, g["seg_loss_grad"], "seg_loss grad")
    check_float(port.seg_refine_by_label(t(g["seg_ps"]), t(g["cls_label"]), 0.01, False), g["refine_masked"], "refine")
    check_float(port.seg_refine_by_label(t(g["seg_ps"]), t(g["cls_label"]), 0.5, True), g["refine_after"], "refine after")
    cp = t(g["cam_pred"]).requires_grad_(True)
    cl = port.cam_loss(cp, t(g["refine_masked"]))
    cl.backward()
    check_float(cl, g["cam_loss"], "cam_loss")
    check_float(cp.grad, g["cam_loss_grad"], "cam_loss grad")


def test_cam_to_label():
    g = load_golden("cam_to_label")
    cam, lab, boxes = t(g["cam"]), t(g["cls_label"]), t(g["boxes"])
    check_labels(port.cam_to_label(cam, lab, bkg_thre=0.5), g["lab_plain"], "plain")
    check_labels(port.cam_to_label(cam, None, bkg_thre=0.5), g["lab_nolabel"], "no cls_label")
    vc, out = port.cam_to_label(cam, lab, img_box=boxes, bkg_thre=0.5, high_thre=0.7, low_thre=0.25,
                                ignore_mid=True, ignore_index=255)
    check_float(vc, g["valid_cam"], "valid_cam")
    check_labels(out, g["lab_box"], "boxed + ignore_mid")
    _, out = port.cam_to_label(cam, lab, img_box=boxes, bkg_thre=0.5, ignore_mid=False, ignore_index=255)
    check_labels(out, g["lab_box_nomid"], "boxed")
    assert out.dtype == torch.int64


@pytest.mark.parametrize("tag", ["a", "b"])
def test_cam2mask(tag):
    g = load_golden("cam2mask_" + tag)
    args = dict(images=t(g["images"]), img_boxes=t(g["boxes"]), cams=t(g["cams"]), cls_labels=t(g["cls_label"]),
                threshold_high=0.7, threshold_low=0.25)
    par = port.ParOracle()
    out = port.cam2mask(**args)
    assert out.dtype == torch.float32
    check_labels(out, g["out_none"], "no refine model")
    check_labels(port.cam2mask(refine_model=par, **args), g["out_par"], "PAR")
    evalbox = dict(args, img_boxes=[[0, -1, 0, -1]] * args["images"].shape[0])
    check_labels(port.cam2mask(refine_model=par, **evalbox), g["out_par_evalbox"], "PAR, eval-style box")
    check_labels(port.cam2mask(downscale=0, **args), g["out_nodownscale"], "downscale=0")
    assert set(np.unique(g["out_par"])) <= set(range(21)) | {255}


def test_energy_function():
    g = load_golden("energy_function")
    segs = t(g["segs"]).requires_grad_(True)
    for filt in ([olat.oracle_bilateralfilter_batch] +
                 ([olat.ref_bilateralfilter_batch] if olat.have_ref() else [])):
        loss, AS, s = port.dense_energy_function_forward(t(g["images"]), segs.detach(), 15, 50.0, t(g["rois"]),
                                                         t(g["unlabel"]), filter_fn=filt)
        check_float(np.array([loss]), g["loss"], "energy loss")
        grad = port.dense_energy_function_backward(torch.tensor([float(g["grad_scale"])]), AS, t(g["rois"]), 2)
        check_float(grad, g["grad_segs"], "energy grad")
        check_float(s, g["filter_in"], "S*ROI")


def test_get_energy_loss():
    g = load_golden("energy_loss")
    logit = t(g["logit"]).requires_grad_(True)
    loss = port.get_energy_loss(t(g["simg"]), logit, t(g["label"]), t(g["boxes"]))
    loss.backward()
    assert loss.shape == (1,)
    check_float(loss, g["loss"], "get_energy_loss")
    check_float(logit.grad, g["grad_logit"], "d loss / d logit")


def test_bilateral_known_answers():
    g = load_golden("bilateral_kat")
    H = W = 224
    out = np.zeros(H * W, np.float32)
    olat.oracle_bilateralfilter_batch(g["kat_img"], np.ones(H * W, np.float32), out, 1, 1, H, W, 15.0, 50.0)
    assert np.array_equal(out.reshape(H, W), g["kat_out"])
    assert abs(out[0] - 87.9367) < 1e-3 and abs(out[H * W // 2 + W // 2] - 324.287) < 1e-2
    _, _, vkeys = olat.oracle_lattice_embed(g["kat_img"], H, W, 15.0, 50.0)
    assert len(vkeys) == int(g["kat_M"]) == 3809
    for tag in ("2x3x5x7", "1x2x9x13", "1x4x24x40"):
        n_, k_, h_, w_ = (int(v) for v in tag.split("x"))
        o = np.zeros(n_ * k_ * h_ * w_, np.float32)
        olat.oracle_bilateralfilter_batch(g["img_" + tag], g["in_" + tag], o, n_, k_, h_, w_, 15.0, 50.0)
        assert np.array_equal(o.reshape(g["out_" + tag].shape), g["out_" + tag]), tag


@pytest.mark.skipif(not olat.have_ref(), reason="oracle/_ref not built (reference tree absent)")
def test_c_oracle_is_bit_exact_against_reference_build():
    rng = np.random.default_rng(3)
    for (N, K, H, W, kind) in ((1, 3, 24, 40, "noise"), (2, 2, 5, 7, "uniform"), (3, 4, 31, 17, "noise"),
                               (1, 21, 56, 56, "uniform")):
        yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
        img = np.stack([127 + 100 * np.sin(0.02 * xx + c) * np.cos(0.03 * yy) for c in range(3)])
        imgs = np.stack([img] * N) + 10 * rng.standard_normal((N, 3, H, W))
        if kind == "uniform":
            imgs = rng.uniform(0, 255, imgs.shape)
        imgs = imgs.astype(np.float32)
        ins = rng.uniform(0, 1, (N, K, H, W)).astype(np.float32)
        a = np.zeros(ins.size, np.float32)
        b = np.zeros(ins.size, np.float32)
        olat.ref_bilateralfilter_batch(imgs, ins, a, N, K, H, W, 15.0, 50.0)
        olat.oracle_bilateralfilter_batch(imgs, ins, b, N, K, H, W, 15.0, 50.0)
        assert np.array_equal(a, b), (N, K, H, W, kind)
