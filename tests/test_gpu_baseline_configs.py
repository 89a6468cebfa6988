"""GPU parity AT the BASELINE.json configs, against the CPU oracle (not only self-consistency).

  configs[1]  VOC  B = 32, 448^2, 21 classes, 2 fg   labels + PAR output on an image slice (every image is
  configs[2]  COCO B = 32, 448^2, 81 classes, 3 fg   independent), CRF loss and gradient
  configs[3]  768^2 and 1024^2, one image each

The GPU always runs the full batch (so the kernels are exercised at the benchmarked shape: 32-image lattice,
energy_*_pair_kernel<81>, the TMA tile path at 384^2 / 512^2 half resolution); the CPU oracle runs the images of a
slice, because cam2mask / PAR are per image and the energy gradient of image b is  -2 AS_b ROI_b / N  with N the batch
size, so  N_gpu * grad_gpu[b] == N_slice * grad_oracle[b].  Every test prints its label-flip count (near-tie protocol,
SURVEY.md 8(d)); run with `pytest -rP` to see them - profiles/ holds the log of the round.
"""
import pytest
import torch
import torch.nn.functional as F

from test_gpu_parity import TOL, assert_close, assert_loss_close, check_labels_near_tie

pytestmark = pytest.mark.gpu
DIL = [1, 2, 4, 8, 12, 24]


@pytest.fixture(scope="module")
def cosa():
    import cosa_b200
    cosa_b200._lib.load()
    return cosa_b200


@pytest.fixture(scope="module")
def port():
    from oracle import reference_port
    return reference_port


def take(host, idx):
    out = {}
    for k, v in host.items():
        out[k] = v[idx].contiguous() if isinstance(v, torch.Tensor) else [v[i] for i in idx]
    return out


def run_config(cosa, port, B, C, H, W, n_fg, seed, thr, slice_idx, what):
    from cosa_b200 import synthetic
    host = synthetic.synthetic_batch(B=B, C=C, H=H, W=W, n_fg=n_fg, seed=seed)
    d = {k: (v.cuda() if k != "img_box" else v) for k, v in host.items()}
    par = cosa.PAR(num_iter=10, dilations=DIL).cuda()
    layer = cosa.DenseEnergyLoss(weight=1e-7, sigma_rgb=15, sigma_xy=100, scale_factor=0.5)
    sl = take(host, slice_idx)
    n_sl = len(slice_idx)
    kw = dict(threshold_high=thr[0], threshold_low=thr[1])

    # ---- the step exactly as bench.py / main.py:117-212 call it, full batch on the GPU -----------------------------
    layer.prebuild_lattice(d["simg"], C)
    label = cosa.cam2mask(images=cosa.denormalize_img(d["simg"]), img_boxes=host["img_box"],
                          cams=cosa.cam_validation(d["cams"], d["cls_label"]), cls_labels=d["cls_label"],
                          refine_model=par, **kw)
    # ---- labels: oracle on the slice, near-tie protocol with the oracle's own margins ------------------------------
    o_img = port.denormalize_img(sl["simg"])
    o_cams = port.cam_validation(sl["cams"], sl["cls_label"])
    want, margins = port.cam2mask(images=o_img, img_boxes=sl["img_box"], cams=o_cams, cls_labels=sl["cls_label"],
                                  refine_model=port.ParOracle(DIL, 10), return_margins=True, **kw)
    flips = check_labels_near_tie(label[slice_idx], want, margins, "%s cam2mask + PAR (derived channel)" % what)
    label_all = cosa.cam2mask(images=cosa.denormalize_img(d["simg"]), img_boxes=host["img_box"],
                              cams=cosa.cam_validation(d["cams"], d["cls_label"]), cls_labels=d["cls_label"],
                              refine_model=par, propagate_all_channels=True, **kw)
    flips_all = check_labels_near_tie(label_all[slice_idx], want, margins, "%s cam2mask + PAR (all channels)" % what)

    # ---- PAR output on the half-resolution stacks of the slice -----------------------------------------------------
    small = F.interpolate(o_img, size=[H // 2, W // 2], mode="bilinear", align_corners=False)
    stack = torch.cat([torch.full((n_sl, 1, H, W), thr[0]), o_cams], dim=1)
    stack = F.interpolate(stack, size=[H // 2, W // 2], mode="bilinear", align_corners=False)
    keys = torch.nonzero(torch.cat([torch.ones(1), sl["cls_label"][0]]))[:, 0]
    masks = stack[:1, keys].softmax(dim=1)
    r_par = assert_close(par(small[:1].cuda(), masks.cuda()), port.par_forward(small[:1], masks, DIL, 10),
                         "%s PAR output" % what)

    # ---- CRF loss and gradient: full batch on the GPU, oracle on the slice -----------------------------------------
    # (both sides are given the ORACLE's labels of the slice so that a near-tie flip cannot leak into the loss;
    #  the other images use the GPU's labels - they only enter through the 1/N of their own gradient rows)
    lab_in = label.clone()
    lab_in[slice_idx] = want.cuda()
    layer.prebuild_lattice(d["simg"], C)
    logit = d["logits"].clone().requires_grad_(True)
    loss = cosa.get_energy_loss(img=d["simg"], logit=logit, label=lab_in, img_box=host["img_box"], loss_layer=layer)
    loss.backward()
    o_logit = sl["logits"].clone().requires_grad_(True)
    o_loss = port.get_energy_loss(sl["simg"], o_logit, want, sl["img_box"])
    o_loss.backward()
    r_grad = assert_close(logit.grad[slice_idx] * (B / n_sl), o_logit.grad, "%s CRF gradient (slice, rescaled)" % what)
    # the loss scalar of the slice alone (a second, small GPU call), and additivity of the big call over slices
    sd = {k: (v.cuda() if k != "img_box" else v) for k, v in sl.items()}
    s_logit = sd["logits"].clone().requires_grad_(True)
    s_loss = cosa.get_energy_loss(img=sd["simg"], logit=s_logit, label=want.cuda(), img_box=sl["img_box"],
                                  loss_layer=layer)
    r_loss, r_loss32, drift = assert_loss_close(port, s_loss, o_loss, "%s CRF loss (slice)" % what)
    total = 0.0
    for i0 in range(0, B, n_sl):
        idx = list(range(i0, min(B, i0 + n_sl)))
        part = take({k: v for k, v in d.items() if k != "img_box"}, idx)
        pl = cosa.get_energy_loss(img=part["simg"], logit=part["logits"], label=lab_in[idx].contiguous(),
                                  img_box=[host["img_box"][i] for i in idx], loss_layer=layer)
        total += float(pl) * len(idx)
    r_add = abs(total / B - float(loss.detach())) / abs(float(loss.detach()))
    assert r_add <= TOL, "%s: loss of the full batch is not the image-weighted mean of its slices (%.3g)" % (what, r_add)
    M = cosa.seg_helper.last_energy_lattice_stats(len(idx), C, H, W)
    assert M[1] == 0
    print("%s: B=%d slice=%s | label flips vs oracle: derived %d, all-channels %d of %d | PAR rel %.2g | "
          "CRF loss rel %.2g vs the oracle accumulated in float64 (%.2g vs its float32 np.dot, whose own drift is %.2g), "
          "grad rel %.2g, additivity %.2g"
          % (what, B, list(slice_idx), flips, flips_all, want.numel(), r_par, r_loss, r_loss32, drift, r_grad, r_add))


def test_voc_b32_448_against_oracle(cosa, port):
    run_config(cosa, port, B=32, C=21, H=448, W=448, n_fg=2, seed=1000, thr=(0.7, 0.25),
               slice_idx=[0, 5, 11, 17, 22, 26, 30, 31], what="VOC configs[1]")


def test_coco_b32_448_against_oracle(cosa, port):
    run_config(cosa, port, B=32, C=81, H=448, W=448, n_fg=3, seed=2000, thr=(0.65, 0.25),
               slice_idx=[0, 9, 18, 31], what="COCO configs[2]")


@pytest.mark.parametrize("size", [768, 1024])
def test_high_resolution_against_oracle(cosa, port, size):
    run_config(cosa, port, B=2, C=21, H=size, W=size, n_fg=2, seed=3000 + size, thr=(0.7, 0.25),
               slice_idx=[1], what="configs[3] %d^2" % size)
