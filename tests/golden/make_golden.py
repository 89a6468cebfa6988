#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/ by running the REFERENCE's own code.

Runs only in the build container, where /root/reference is mounted (read-only).  The reference is
imported as is:
  * models/PAR.py is loaded by file path (importing ``models`` would pull in timm, absent here);
  * utils/seg_helper.py is imported with a ctypes shim module named ``bilateralfilter`` in front of
    the UNMODIFIED reference C++ (oracle/_ref/libbf_ref.so, built by oracle/Makefile) in place of
    the SWIG glue, and with ``pydensecrf`` stubbed (not on this path);
  * ``Tensor.cuda`` is neutralised, because this container has no GPU (seg_helper.py:230,880,901).
Nothing else is patched.  Inputs are seeded; re-running reproduces the committed files bit for bit on
the same torch build.

    python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import lattice as olat  # noqa: E402
from cosa_b200 import synthetic as port  # noqa: E402  (the shared synthetic-input generator)


def load_reference():
    """The reference's own modules (oracle/reference_import.py holds the import shim shared with bench.py)."""
    from oracle import reference_import
    global ref_torch_helper
    par, sh, ref_torch_helper = reference_import.load_reference()
    return par, sh, None


ONLY = set(sys.argv[1:])        # optional: names of the fixtures to (re)write; default all


def save(name, **arrays):
    if ONLY and name not in ONLY:
        return
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


def main():
    par_mod, sh, _ = load_reference()
    torch.manual_seed(0)
    meta = dict(torch=torch.__version__, numpy=np.__version__, cpu_capability=torch.backends.cpu.get_cpu_capability())

    # ---- PAR ---------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(11)
    imgs = torch.randint(0, 256, (2, 3, 40, 56), generator=g).float() / 255.0
    imgs[1] = (imgs[1] * 0.1 + 0.4)                     # low-contrast image: small std48
    imgs[1, :, :8, :8] = 0.5                            # constant patch: std48 = 0 branch
    masks = torch.rand((2, 4, 40, 56), generator=g).softmax(dim=1)
    par = par_mod.PAR(num_iter=10, dilations=[1, 2, 4, 8, 12, 24])
    out = par(imgs, masks)
    par3 = par_mod.PAR(num_iter=3, dilations=[1, 2, 4])
    out3 = par3(imgs, masks)
    masks_lr = torch.rand((2, 3, 20, 28), generator=g)  # resized to the image size with align_corners=True
    out_lr = par(imgs, masks_lr)
    save("par", imgs=imgs, masks=masks, out=out, out_iter3_dil124=out3, masks_lr=masks_lr, out_lr=out_lr,
         pos_softmax=torch.softmax(-(par.pos.flatten() / (par.pos.flatten().std() + 1e-8) / 0.3) ** 2, 0))

    # ---- CAM normalise (seg_helper.py:264-270 on three pre-made scale maps) --------------------
    scales = [torch.rand((2, 5, 24, 32), generator=g) * s for s in (1.0, 0.7, 1.3)]
    cam = torch.sum(torch.stack(scales, dim=0), dim=0)
    cam = cam + torch.nn.functional.adaptive_max_pool2d(-cam, (1, 1))
    cam /= torch.nn.functional.adaptive_max_pool2d(cam, (1, 1)) + 1e-5
    save("normalize", s0=scales[0], s1=scales[1], s2=scales[2], out=cam)

    # ---- multi_scale_camseg with a stub teacher (seg_helper.py:232-275) ----------------------------------
    class StubTeacher(torch.nn.Module):
        """Deterministic stand-in for VITNetwork.forward: 16x16 patch means through fixed 1x1 mixes."""

        def __init__(self):
            super().__init__()
            gg = torch.Generator().manual_seed(77)
            self.w_cam = torch.randn((5, 3), generator=gg)
            self.w_aux = torch.randn((5, 3), generator=gg)
            self.w_seg = torch.randn((6, 3), generator=gg)
            self.calls = []

        def forward(self, x, cam_only=False):
            tok = torch.nn.functional.avg_pool2d(x, 16)
            mix = lambda wgt: torch.einsum("oc,bchw->bohw", wgt, tok) + 0.1 * torch.sin(7.0 * tok.sum(1, keepdim=True))
            out = (None, None, None, mix(self.w_seg), mix(self.w_cam), mix(self.w_aux))
            self.calls.append(out[3:])
            return out

    teacher = StubTeacher()
    wimg = torch.randn((2, 3, 64, 96), generator=g)
    m_cam, m_aux, m_seg = sh.multi_scale_camseg(teacher, wimg, [1.0, 0.5, 1.5])
    raw = {}
    for i, (sg, cm, ax) in enumerate(teacher.calls):
        raw["raw_seg%d" % i], raw["raw_cam%d" % i], raw["raw_aux%d" % i] = sg, cm, ax
    save("multi_scale", imgs=wimg, cam=m_cam, cam_aux=m_aux, seg=m_seg, **raw)

    # ---- denormalize_img (utils/torch_helper.py:354-367, main.py:117) ---------------------------
    dn = port.synthetic_batch(B=2, C=4, H=24, W=40, n_fg=1, seed=21)
    x = dn["simg"].clone()
    x[1, :, :4] += torch.rand((3, 4, 40), generator=g) * 0.01      # values that are not exact (u8 - mean) / std
    mean_, std_ = torch.tensor(port.IMAGENET_MEAN).view(3, 1), torch.tensor(port.IMAGENET_STD).view(3, 1)
    ramp = torch.arange(0, 256, dtype=torch.float32).view(1, 256)   # every uint8 level in every channel
    x[0].view(3, -1)[:, :256] = (ramp - mean_) / std_
    save("denormalize", simg=x, out=ref_torch_helper.denormalize_img(x))

    # ---- cam_validation / cam_to_label ---------------------------------------------------------
    d = port.synthetic_batch(B=3, C=6, H=48, W=64, n_fg=2, seed=5, cam_kind="grid")
    raw_cam = torch.rand((3, 5, 48, 64), generator=g)
    valid = sh.cam_validation(raw_cam, d["cls_label"])
    lab_plain = sh.cam_to_label(raw_cam.clone(), d["cls_label"], bkg_thre=0.5)
    lab_nolabel = sh.cam_to_label(raw_cam.clone(), None, bkg_thre=0.5)
    boxes = torch.tensor([[0, 48, 0, 64], [4, 40, 8, 60], [0, -1, 0, -1]], dtype=torch.int16)
    vc, lab_box = sh.cam_to_label(raw_cam.clone(), d["cls_label"], img_box=boxes, bkg_thre=0.5, high_thre=0.7,
                                  low_thre=0.25, ignore_mid=True, ignore_index=255)
    vc2, lab_box_nomid = sh.cam_to_label(raw_cam.clone(), d["cls_label"], img_box=boxes, bkg_thre=0.5,
                                         ignore_mid=False, ignore_index=255)
    save("cam_to_label", cam=raw_cam, cls_label=d["cls_label"], boxes=boxes, valid=valid, lab_plain=lab_plain,
         lab_nolabel=lab_nolabel, valid_cam=vc, lab_box=lab_box, lab_box_nomid=lab_box_nomid)

    # ---- cam2mask: no refine model (the shipped default) and with PAR --------------------------
    for tag, kw in (("a", dict(B=3, C=6, H=64, W=96, n_fg=2, seed=21, cam_kind="blobs", box="crop")),
                    ("b", dict(B=2, C=21, H=96, W=64, n_fg=3, seed=22, cam_kind="grid", box="full"))):
        d = port.synthetic_batch(**kw)
        if tag == "b":
            d["cls_label"][1] = 0                      # an image with no foreground class (nc = 1)
            d["cams"][1] = 0
        args = dict(images=d["img_denorm"], img_boxes=d["img_box"], cams=d["cams"], cls_labels=d["cls_label"],
                    threshold_high=0.7, threshold_low=0.25)
        m_none = sh.cam2mask(**args)
        m_par = sh.cam2mask(refine_model=par, **args)
        m_list = sh.cam2mask(**dict(args, img_boxes=[[0, -1, 0, -1]] * kw["B"]), refine_model=par)
        m_nods = sh.cam2mask(**args, downscale=0)
        save("cam2mask_" + tag, images=d["img_denorm"], boxes=d["img_box"], cams=d["cams"],
             cls_label=d["cls_label"], out_none=m_none, out_par=m_par, out_par_evalbox=m_list, out_nodownscale=m_nods)

    # ---- dense-CRF energy: the autograd Function alone, the layer, and get_energy_loss ---------
    d = port.synthetic_batch(B=2, C=6, H=64, W=96, n_fg=2, seed=31, box="crop")
    label = sh.cam2mask(images=d["img_denorm"], img_boxes=d["img_box"], cams=d["cams"], cls_labels=d["cls_label"],
                        threshold_high=0.7, threshold_low=0.25)
    layer = sh.DenseEnergyLoss(weight=1e-7, sigma_rgb=15, sigma_xy=100, scale_factor=0.5)
    logit = d["logits"].clone().requires_grad_(True)
    loss = sh.get_energy_loss(img=d["simg"], logit=logit, label=label, img_box=d["img_box"], loss_layer=layer)
    loss.backward()
    save("energy_loss", simg=d["simg"], logit=d["logits"], label=label, boxes=d["img_box"], loss=loss,
         grad_logit=logit.grad)

    g2 = torch.Generator().manual_seed(41)
    N, K, H, W = 2, 5, 32, 48
    images = torch.randint(0, 256, (N, 3, H, W), generator=g2).float()
    images[1] = d["simg"][1, :, :H, :W] * 58.0 + 120.0
    segs = torch.rand((N, K, H, W), generator=g2).softmax(dim=1).requires_grad_(True)
    rois = torch.zeros((N, H, W))
    rois[0, 2:30, 4:44] = 1
    rois[1] = 1
    unlabel = torch.rand((N, H, W), generator=g2) < 0.2
    fl = sh.DenseEnergyLossFunction.apply(images, segs, 15, 50.0, rois.clone(), unlabel)
    (fl * 3.0).sum().backward()
    # the gated filter response saved for backward is recoverable from the gradient; store the raw
    # filter response too (reference C++ through the shim), for the bilateralfilter_batch drop-in.
    s_roi = (segs.detach() * rois[:, None]).contiguous()
    AS = np.zeros(s_roi.numel(), np.float32)
    olat.ref_bilateralfilter_batch(images.numpy().reshape(-1), s_roi.numpy().reshape(-1), AS, N, K, H, W, 15.0, 50.0)
    save("energy_function", images=images, segs=segs, rois=rois, unlabel=unlabel, loss=fl, grad_segs=segs.grad,
         grad_scale=np.float32(3.0), filter_in=s_roi, filter_out=AS.reshape(N, K, H, W))

    # ---- bilateral filter known-answer tests ---------------------------------------------------
    # (1) SURVEY.md 8(c) smoke KAT: analytic image, all-ones plane, 224^2.
    H = W = 224
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    kat_img = np.stack([127 + 100 * np.sin(np.float32(0.02) * xx + np.float32(c)) * np.cos(np.float32(0.03) * yy)
                        for c in range(3)]).astype(np.float32)
    ones = np.ones((1, 1, H, W), np.float32)
    kat_out = np.zeros(H * W, np.float32)
    olat.ref_bilateralfilter_batch(kat_img.reshape(-1), ones.reshape(-1), kat_out, 1, 1, H, W, 15.0, 50.0)
    # (2) ragged sizes: H*W not a multiple of 4 exercises the SSE padding pixels (permutohedral.cpp:168-173).
    rng = np.random.default_rng(7)
    small = {}
    for (n_, k_, h_, w_) in ((2, 3, 5, 7), (1, 2, 9, 13), (1, 4, 24, 40)):
        im = rng.uniform(0, 255, (n_, 3, h_, w_)).astype(np.float32)
        xin = rng.uniform(0, 1, (n_, k_, h_, w_)).astype(np.float32)
        o = np.zeros(xin.size, np.float32)
        olat.ref_bilateralfilter_batch(im.reshape(-1), xin.reshape(-1), o, n_, k_, h_, w_, 15.0, 50.0)
        tag = "%dx%dx%dx%d" % (n_, k_, h_, w_)
        small["img_" + tag], small["in_" + tag], small["out_" + tag] = im, xin, o.reshape(xin.shape)
    save("bilateral_kat", kat_img=kat_img,
         kat_out=kat_out.reshape(H, W), kat_M=np.int32(3809), **small)

    # ---- seg_loss, seg_refine_by_label, cam_loss -----------------------------------
    g = torch.Generator().manual_seed(23)
    d = port.synthetic_batch(B=2, C=21, H=48, W=64, n_fg=2, seed=9)
    lbl = torch.randint(0, 21, (2, 48, 64), generator=g).float()
    lbl[torch.rand((2, 48, 64), generator=g) < 0.3] = 255
    lbl[torch.rand((2, 48, 64), generator=g) < 0.4] = 0
    sp = d["logits"].clone().requires_grad_(True)
    sl = sh.seg_loss(sp, lbl, fg_alpha=0.5)
    sl.backward()
    sp2 = d["logits"].clone().requires_grad_(True)
    sl2 = sh.seg_loss(sp2, torch.full_like(lbl, 255), fg_alpha=0.3)     # nothing labelled: both terms 0 / 1e-6
    sl2.backward()
    seg_ps = 2.0 * d["logits"]
    v_a = sh.seg_refine_by_label(seg_ps, d["cls_label"], softmaxtemp=0.01, after_softmax=False)
    v_b = sh.seg_refine_by_label(seg_ps, d["cls_label"], softmaxtemp=0.5, after_softmax=True)
    cam_pred = torch.randn((2, 20, 6, 8), generator=g).requires_grad_(True)
    cl = sh.cam_loss(cam_pred, v_a)
    cl.backward()
    cam_pred2 = cam_pred.detach().clone().requires_grad_(True)
    cl2 = sh.cam_loss(cam_pred2, v_b, is_relu=False)
    cl2.backward()
    save("losses", logits=d["logits"], label=lbl, cls_label=d["cls_label"], seg_loss=sl, seg_loss_grad=sp.grad,
         seg_loss_empty=sl2, seg_loss_empty_grad=sp2.grad, seg_ps=seg_ps, refine_masked=v_a, refine_after=v_b,
         cam_pred=cam_pred, cam_loss=cl, cam_loss_grad=cam_pred.grad, cam_loss_norelu=cl2,
         cam_loss_norelu_grad=cam_pred2.grad)

    # ---- dense-CRF inference (seg_helper.py:961-996): pydensecrf is absent, so what the reference CAN pin here is its
    # own Permutohedral class (the code base pydensecrf wraps) for the two filters of the mean-field update, driven
    # through oracle/ref_lattice_shim.cpp: filter responses on the 2-D and 5-D lattices, and the marginals of the
    # update rule restated in oracle/crf_oracle.py evaluated WITH the reference class as the filter ------------------
    from oracle import crf_oracle as co
    rng = np.random.default_rng(2024)
    Hc, Wc, Cc = 45, 61, 6                                  # H*W % 4 != 0: the SSE padding points take part
    yy, xx = np.mgrid[0:Hc, 0:Wc].astype(np.float32)
    base = np.stack([127 + 100 * np.sin(0.05 * xx + c) * np.cos(0.07 * yy) for c in range(3)], -1)
    crf_img = np.clip(np.floor(base + 12 * rng.standard_normal((Hc, Wc, 3))), 0, 255).astype(np.uint8)
    logits_c = 2.5 * rng.standard_normal((Cc, Hc // 4 + 1, Wc // 4 + 1)).astype(np.float32)
    logits_c = torch.nn.functional.interpolate(torch.from_numpy(logits_c)[None], size=(Hc, Wc), mode="bilinear",
                                               align_corners=False)[0]
    crf_probs = logits_c.softmax(dim=0).numpy().astype(np.float32)
    vals = rng.random((3, Hc * Wc)).astype(np.float32)
    f_g1, f_g4 = co.gaussian_features(Hc, Wc, 1.0), co.gaussian_features(Hc, Wc, 4.0)
    f_b = co.bilateral_features(crf_img, 121, 5)
    save("crf_inference", image=crf_img, probs=crf_probs, vals=vals,
         filt_gauss1=co.ref_filter(f_g1, vals), filt_gauss4=co.ref_filter(f_g4, vals), filt_bilateral=co.ref_filter(f_b, vals),
         q_infv2=co.crf_inference(crf_img, crf_probs, 1, 1, 1, 4, 121, 5, filter_fn=co.ref_filter),
         q_inf_t10=co.crf_inference(crf_img, crf_probs, 10, 3, 4, 3, 83, 5, filter_fn=co.ref_filter),
         q_iter0=co.crf_inference(crf_img, crf_probs, 0, 1, 1, 4, 121, 5, filter_fn=co.ref_filter))

    with open(os.path.join(HERE, "META.txt"), "w") as f:
        for k, v in meta.items():
            f.write("%s: %s\n" % (k, v))


if __name__ == "__main__":
    main()
