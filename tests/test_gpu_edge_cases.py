"""GPU parity on the edge cases of the path: ragged / odd geometries (generic kernels), COCO-sized class counts,
empty and missing boxes, images without foreground, batches above the 64-image lattice chunk, exotic dilations,
and the lattice key-range guard.  Same bars as tests/test_gpu_parity.py."""
import numpy as np
import pytest
import torch

from conftest import rel_inf
from test_gpu_parity import NEAR_TIE, TOL, _oracle_margin, assert_close, assert_same, check_labels_near_tie

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


def batch(**kw):
    from cosa_b200 import synthetic
    return synthetic.synthetic_batch(**kw)


def to_cuda(d):
    return {k: (v.cuda() if k != "img_box" else v) for k, v in d.items()}


@pytest.mark.parametrize("kw", [
    dict(B=2, C=81, H=96, W=96, n_fg=3, seed=101),                   # COCO class count
    dict(B=2, C=21, H=66, W=90, n_fg=2, seed=102, box="crop"),       # half-res 33x45: generic (non-vector) kernels
    dict(B=1, C=21, H=50, W=72, n_fg=5, seed=103),                   # 25x36: vector path with nch = 12 (two passes)
    dict(B=3, C=4, H=32, W=48, n_fg=3, seed=104, cam_kind="grid"),   # all classes present
])
def test_cam2mask_and_energy_vs_oracle(cosa, port, kw):
    host = batch(**kw)
    d = to_cuda(host)
    par = cosa.PAR(num_iter=10, dilations=DIL).cuda()
    args = dict(img_boxes=host["img_box"], threshold_high=0.7, threshold_low=0.25)
    got = cosa.cam2mask(images=d["img_denorm"], cams=d["cams"], cls_labels=d["cls_label"], refine_model=par, **args)
    want = port.cam2mask(images=host["img_denorm"], cams=host["cams"], cls_labels=host["cls_label"],
                         refine_model=port.ParOracle(), **args)
    margins = _oracle_margin(port, dict(images=host["img_denorm"], cams=host["cams"], cls_label=host["cls_label"]),
                             port.ParOracle())
    n = check_labels_near_tie(got, want, margins, "cam2mask %s" % kw)
    # energy loss on the ORACLE's labels so that both sides see identical inputs
    layer = cosa.DenseEnergyLoss(weight=1e-7, sigma_rgb=15, sigma_xy=100, scale_factor=0.5)
    logit = d["logits"].clone().requires_grad_(True)
    loss = cosa.get_energy_loss(img=d["simg"], logit=logit, label=want.cuda(), img_box=host["img_box"], loss_layer=layer)
    loss.backward()
    o_logit = host["logits"].clone().requires_grad_(True)
    o_loss = port.get_energy_loss(host["simg"], o_logit, want, host["img_box"])
    o_loss.backward()
    r1 = assert_close(loss, o_loss, "energy loss %s" % kw)
    r2 = assert_close(logit.grad, o_logit.grad, "energy grad %s" % kw)
    print("%s: label flips %d, loss rel %.2g, grad rel %.2g" % (kw, n, r1, r2))


def test_boxes_empty_missing_and_no_foreground(cosa, port):
    host = batch(B=4, C=6, H=48, W=64, n_fg=2, seed=7)
    host["cls_label"][2] = 0                          # image 2: no foreground class -> nc = 1, all background
    host["cams"][2] = 0
    d = to_cuda(host)
    boxes = [[0, 48, 0, 64], [10, 10, 0, 64], [5, 40, -20, -4]]      # full, empty rows, negative ends; image 3: no box
    par = cosa.PAR(num_iter=10, dilations=DIL).cuda()
    got = cosa.cam2mask(images=d["img_denorm"], img_boxes=boxes, cams=d["cams"], cls_labels=d["cls_label"],
                        threshold_high=0.7, threshold_low=0.25, refine_model=par)
    want = port.cam2mask(images=host["img_denorm"], img_boxes=boxes, cams=host["cams"], cls_labels=host["cls_label"],
                         threshold_high=0.7, threshold_low=0.25, refine_model=port.ParOracle())
    margins = _oracle_margin(port, dict(images=host["img_denorm"], cams=host["cams"], cls_label=host["cls_label"]),
                             port.ParOracle())
    check_labels_near_tie(got, want, margins, "cam2mask with empty / missing / negative boxes")
    assert bool((got[1] == 255).all()) and bool((got[3] == 255).all())          # empty box / no box: all ignore
    assert set(got[2].unique().tolist()) <= {0.0, 255.0}
    # the same boxes through cam_to_label and get_energy_loss
    _, lab = cosa.cam_to_label(d["cams"], d["cls_label"], img_box=boxes, bkg_thre=0.5, high_thre=0.7, low_thre=0.25,
                               ignore_mid=True, ignore_index=255)
    _, olab = port.cam_to_label(host["cams"], host["cls_label"], img_box=boxes, bkg_thre=0.5, high_thre=0.7,
                                low_thre=0.25, ignore_mid=True, ignore_index=255)
    assert_same(lab, olab, "cam_to_label with ragged boxes")
    layer = cosa.DenseEnergyLoss(weight=1e-7, sigma_rgb=15, sigma_xy=100, scale_factor=0.5)
    logit = d["logits"].clone().requires_grad_(True)
    loss = cosa.get_energy_loss(img=d["simg"], logit=logit, label=want.cuda(), img_box=boxes, loss_layer=layer)
    loss.backward()
    o_logit = host["logits"].clone().requires_grad_(True)
    o_loss = port.get_energy_loss(host["simg"], o_logit, want, boxes)
    o_loss.backward()
    assert_close(loss, o_loss, "energy loss, ragged boxes")
    assert_close(logit.grad, o_logit.grad, "energy grad, ragged boxes")


def test_energy_function_more_than_one_lattice_chunk(cosa, port):
    """N = 70 > 64 images: the lattice runs in two chunks that share one loss accumulator."""
    gen = torch.Generator().manual_seed(5)
    N, K, H, W = 70, 3, 10, 12
    images = torch.randint(0, 256, (N, 3, H, W), generator=gen).float()
    segs = torch.rand((N, K, H, W), generator=gen).softmax(dim=1)
    rois = (torch.rand((N, H, W), generator=gen) < 0.9).float()
    unlabel = torch.rand((N, H, W), generator=gen) < 0.2
    want_loss, want_as, _ = port.dense_energy_function_forward(images, segs, 15.0, 50.0, rois, unlabel)
    s = segs.cuda().requires_grad_(True)
    loss = cosa.DenseEnergyLossFunction.apply(images.cuda(), s, 15.0, 50.0, rois.cuda(), unlabel.cuda())
    loss.backward()
    assert_close(loss, np.array([want_loss]), "chunked energy loss")
    want_grad = port.dense_energy_function_backward(torch.ones(1), want_as, rois, N)
    assert_close(s.grad, want_grad, "chunked energy grad")


@pytest.mark.parametrize("dilations", [[3, 6], [1, 5, 9, 32], [2], [1, 2, 4, 8, 12, 24, 30]])
def test_par_exotic_dilations(cosa, port, dilations):
    """Unaligned dilations (scalar branch of the vector kernel), pads wider than 24 (plain layout), 1 and 7 dilations."""
    gen = torch.Generator().manual_seed(sum(dilations))
    imgs = torch.randint(0, 256, (2, 3, 36, 48), generator=gen).float() / 255
    masks = torch.rand((2, 5, 36, 48), generator=gen)
    got = cosa.PAR(num_iter=4, dilations=dilations)(imgs.cuda(), masks.cuda())
    assert_close(got, port.par_forward(imgs, masks, tuple(dilations), 4), "PAR dilations %s" % dilations)


def test_par_zero_iterations_and_resize_only(cosa, port):
    gen = torch.Generator().manual_seed(1)
    imgs = torch.rand((1, 3, 24, 32), generator=gen)
    masks = torch.rand((1, 3, 12, 16), generator=gen)
    assert_close(cosa.PAR(num_iter=0, dilations=DIL)(imgs.cuda(), masks.cuda()),
                 port.par_forward(imgs, masks, DIL, 0), "PAR with num_iter=0 (resize only)")
    same = torch.rand((1, 3, 24, 32), generator=gen)
    assert torch.equal(cosa.PAR(num_iter=0, dilations=DIL)(imgs.cuda(), same.cuda()).cpu(), same)


def test_cam_normalize_large_planes(cosa, port):
    gen = torch.Generator().manual_seed(3)
    scales = [torch.rand((2, 20, 448, 448), generator=gen) * s - 0.1 for s in (1.0, 0.5, 1.5)]
    got = cosa.cam_normalize([s.cuda() for s in scales])
    assert_same(got, port.normalize_cam(scales), "cam_normalize 448x448 (multi-CTA min/max)")
    one = cosa.cam_normalize(scales[0].cuda())
    assert_same(one, port.normalize_cam([scales[0]]), "cam_normalize, single tensor")


def test_lattice_key_range_guard(cosa):
    """Coordinates beyond the packed-key range must raise the flag instead of aliasing silently."""
    from cosa_b200 import bilateralfilter as bf
    N, K, H, W = 1, 2, 16, 16
    img = torch.rand((N, 3, H, W), device="cuda") * 255
    x = torch.rand((N, K, H, W), device="cuda")
    out = torch.empty_like(x)
    bf.bilateralfilter_batch(img, x, out, N, K, H, W, 15.0, 50.0)
    assert bf.lattice_stats(N, K, H, W)[1] == 0
    assert bool(torch.isfinite(out).all())
    bf.bilateralfilter_batch(img, x, out, N, K, H, W, 0.001, 50.0)       # sigma_rgb = 0.001 -> coordinates ~ 1e6
    assert bf.lattice_stats(N, K, H, W)[1] == 1
    assert bool(torch.isnan(out).all()), "an out-of-range lattice must poison the output, not alias vertices"
    # ... the dense-CRF loss built on it is NaN as well (fail loudly on the production path) ...
    segs = torch.rand((N, K, H, W), device="cuda").softmax(dim=1)
    loss = cosa.DenseEnergyLossFunction.apply(img, segs, 0.001, 50.0, torch.ones((N, H, W), device="cuda"),
                                              torch.zeros((N, H, W), dtype=torch.bool, device="cuda"))
    assert bool(torch.isnan(loss).all())
    # ... and the synchronous host form (the SWIG drop-in) reports it as an error
    import numpy as np
    outs = np.zeros(N * K * H * W, dtype=np.float32)
    with pytest.raises(cosa._lib.CosaError):
        bf.bilateralfilter_batch(img.cpu().numpy().ravel(), x.cpu().numpy().ravel(), outs, N, K, H, W, 0.001, 50.0)
    # the reference's own range (short) is wider: sigma_rgb = 0.05 is fine there and flagged here, 0.5 is fine in both
    bf.bilateralfilter_batch(img, x, out, N, K, H, W, 0.5, 50.0)
    assert bf.lattice_stats(N, K, H, W)[1] == 0 and bool(torch.isfinite(out).all())


def test_high_resolution_energy_matches_oracle(cosa, port):
    """One image of the 512^2 sweep point (BASELINE.json configs[3]) against the oracle."""
    host = batch(B=1, C=21, H=512, W=512, n_fg=2, seed=512)
    d = to_cuda(host)
    label = port.cam2mask(images=host["img_denorm"], img_boxes=host["img_box"], cams=host["cams"],
                          cls_labels=host["cls_label"], threshold_high=0.7, threshold_low=0.25)
    layer = cosa.DenseEnergyLoss(weight=1e-7, sigma_rgb=15, sigma_xy=100, scale_factor=0.5)
    logit = d["logits"].clone().requires_grad_(True)
    loss = cosa.get_energy_loss(img=d["simg"], logit=logit, label=label.cuda(), img_box=host["img_box"], loss_layer=layer)
    loss.backward()
    o_logit = host["logits"].clone().requires_grad_(True)
    o_loss = port.get_energy_loss(host["simg"], o_logit, label, host["img_box"])
    o_loss.backward()
    assert_close(loss, o_loss, "512^2 energy loss")
    assert_close(logit.grad, o_logit.grad, "512^2 energy grad")
    got = cosa.cam2mask(images=d["img_denorm"], img_boxes=host["img_box"], cams=d["cams"], cls_labels=d["cls_label"],
                        threshold_high=0.7, threshold_low=0.25)
    margins = _oracle_margin(port, dict(images=host["img_denorm"], cams=host["cams"], cls_label=host["cls_label"]), None)
    print("512^2 cam2mask (no refine model): %d label flips" % check_labels_near_tie(got, label, margins, "512^2 cam2mask"))


def test_host_pipeline_matches_device_path(cosa, port):
    """cosa_b200.HostPipeline (pinned host buffers in, labels + loss out, sparse CAM upload, two staging slots)
    returns exactly what the device-resident calls return, batch after batch."""
    dil = [1, 2, 4, 8, 12, 24]
    par = cosa.PAR(dil, 10).cuda()
    layer = cosa.DenseEnergyLoss(weight=1e-7, sigma_rgb=15, sigma_xy=100, scale_factor=0.5)
    pipe = cosa.HostPipeline(par, layer, threshold_high=0.7, threshold_low=0.25, want_grad=True)
    batches = [batch(B=3, C=21, H=64, W=96, n_fg=2, seed=s) for s in (1, 2, 3)]
    outs = []
    for hb in batches:
        # garbage in the planes of absent classes must not matter (they are never uploaded)
        hb = dict(hb)
        hb["cams"] = hb["cams"] + 7.0 * (1 - hb["cls_label"])[:, :, None, None]
        pinned = {k: (v.pin_memory() if k != "img_box" else v) for k, v in hb.items()}
        r = pipe.submit(pinned)
        if r is not None:
            outs.append(tuple(t.clone() for t in r))
    outs += [tuple(t.clone() for t in r) for r in pipe.drain()]
    assert len(outs) == 3
    for hb, (label, loss, grad) in zip(batches, outs):
        d = to_cuda(hb)
        cams = cosa.cam_validation(d["cams"], d["cls_label"])
        want = cosa.cam2mask(images=cosa.denormalize_img(d["simg"]), img_boxes=hb["img_box"], cams=cams,
                             cls_labels=d["cls_label"], threshold_high=0.7, threshold_low=0.25, refine_model=par)
        logit = d["logits"].clone().requires_grad_(True)
        wl = cosa.get_energy_loss(img=d["simg"], logit=logit, label=want, img_box=hb["img_box"], loss_layer=layer)
        wl.backward()
        assert torch.equal(label, want.cpu())
        assert abs(float(loss) - float(wl.detach())) <= 1e-6 * abs(float(wl.detach()))
        assert float((grad - logit.grad.cpu()).abs().max()) <= 1e-5 * float(logit.grad.abs().max())
    assert pipe.h2d_bytes < 3 * sum(v.numel() * 4 for k, v in batches[0].items() if k != "img_box")
    with pytest.raises(ValueError):
        pipe.submit({k: (v.cuda() if k != "img_box" else v) for k, v in batches[0].items()})


def test_cam_merge_with_labels_equals_merge_then_validation(cosa):
    """multi_scale_cam_merge(..., cls_label) == cam_validation(multi_scale_cam_merge(...)) (main.py:135-137), bit for
    bit, for 0/1 labels, for a fractional label, and on a width the row-walking kernel does not take."""
    gen = torch.Generator().manual_seed(3)
    for (B, C1, H, W) in ((3, 20, 64, 96), (2, 5, 30, 41)):
        raw = [torch.randn((2 * B, C1, g_, g_ + 1), generator=gen).cuda() for g_ in (4, 2, 6)]
        cls = (torch.rand((B, C1), generator=gen) < 0.2).float()
        cls[0, 0] = 0.5
        want = cosa.cam_validation(cosa.multi_scale_cam_merge(raw, (H, W)), cls.cuda())
        got = cosa.multi_scale_cam_merge(raw, (H, W), cls_label=cls.cuda())
        assert torch.equal(got, want), (B, C1, H, W)


def test_host_pipeline_native_inputs_match_device_path(cosa):
    """HostPipeline.submit_native (raw multi-scale CAMs + token-grid logits in pinned host memory) against the same
    steps called one by one on device tensors: labels identical, loss and token-grid gradient equal."""
    from cosa_b200 import synthetic
    par = cosa.PAR(num_iter=10, dilations=DIL).cuda()
    layer = cosa.DenseEnergyLoss(weight=1e-7, sigma_rgb=15, sigma_xy=100, scale_factor=0.5)
    pipe = cosa.HostPipeline(par, layer, threshold_high=0.7, threshold_low=0.25, want_grad=True)
    batches = []
    for i in range(3):
        hb = batch(B=2, C=21, H=96, W=128, n_fg=2, seed=300 + i)
        hb["raw_cams"] = synthetic.synthetic_raw_cams(hb, seed=400 + i)
        batches.append(hb)
    outs = []
    for hb in batches:
        nb = dict(simg=hb["simg"].pin_memory(), raw_cams=[t.pin_memory() for t in hb["raw_cams"]],
                  seg_lowres=hb["seg_lowres"].pin_memory(), cls_label=hb["cls_label"].pin_memory(), img_box=hb["img_box"])
        r = pipe.submit_native(nb)
        if r is not None:
            outs.append(tuple(t.clone() for t in r))
    outs += [tuple(t.clone() for t in r) for r in pipe.drain()]
    assert len(outs) == 3
    for hb, (label, loss, grad) in zip(batches, outs):
        H, W = hb["simg"].shape[2:]
        cams = cosa.cam_validation(cosa.multi_scale_cam_merge([t.cuda() for t in hb["raw_cams"]], (H, W)),
                                   hb["cls_label"].cuda())
        want = cosa.cam2mask(images=cosa.denormalize_img(hb["simg"].cuda()), img_boxes=hb["img_box"], cams=cams,
                             cls_labels=hb["cls_label"].cuda(), threshold_high=0.7, threshold_low=0.25,
                             refine_model=par)
        low = hb["seg_lowres"].cuda().requires_grad_(True)
        wl = cosa.get_energy_loss(img=hb["simg"].cuda(), logit=cosa.upsample_bilinear(low, (H, W)), label=want,
                                  img_box=hb["img_box"], loss_layer=layer)
        wl.backward()
        assert torch.equal(label, want.cpu())
        assert_close(loss, wl, "native pipeline loss", tol=1e-5)
        assert_close(grad, low.grad, "native pipeline token-grid gradient", tol=1e-5)
    assert pipe.h2d_bytes < 0.3 * 3 * sum(batches[0][k].numel() * 4 for k in ("simg", "cams", "logits"))


def test_shared_affinity_between_cam2mask_calls(cosa):
    """main.py:158 and :191 call cam2mask twice per batch (CAMs, auxiliary CAMs) on the same images: inside
    PAR.shared_affinity() the second call reuses the first call's affinity.  Labels must equal the ones of two
    independent calls; a call on other images / another geometry inside the context recomputes."""
    from cosa_b200 import _lib
    par = cosa.PAR(num_iter=10, dilations=DIL).cuda()
    d = to_cuda(batch(B=3, C=21, H=128, W=160, n_fg=2, seed=51))
    aux = to_cuda(batch(B=3, C=21, H=128, W=160, n_fg=2, seed=52))
    aux_cams = aux["cams"].flip(-1) * 0.9 + 0.05 * d["cams"]
    kw = dict(img_boxes=d["img_box"], cls_labels=d["cls_label"], threshold_high=0.7, threshold_low=0.25, refine_model=par)
    want1 = cosa.cam2mask(images=d["img_denorm"], cams=d["cams"], **kw)
    want2 = cosa.cam2mask(images=d["img_denorm"], cams=aux_cams, **kw)
    small = to_cuda(batch(B=2, C=21, H=64, W=96, n_fg=2, seed=53))
    want3 = cosa.cam2mask(images=small["img_denorm"], cams=small["cams"], img_boxes=small["img_box"],
                          cls_labels=small["cls_label"], threshold_high=0.7, threshold_low=0.25, refine_model=par)
    n0 = _lib.launch_count()
    with par.shared_affinity():
        got1 = cosa.cam2mask(images=d["img_denorm"], cams=d["cams"], **kw)
        n1 = _lib.launch_count()
        got2 = cosa.cam2mask(images=d["img_denorm"], cams=aux_cams, **kw)
        n2 = _lib.launch_count()
        got3 = cosa.cam2mask(images=small["img_denorm"], cams=small["cams"], img_boxes=small["img_box"],
                             cls_labels=small["cls_label"], threshold_high=0.7, threshold_low=0.25, refine_model=par)
        # an unrelated user of the scratch buffer in between invalidates the shared affinity
        got1b = cosa.cam2mask(images=d["img_denorm"], cams=d["cams"], **kw)
        cosa.PAR(num_iter=1, dilations=DIL).cuda()(small["img_denorm"], small["cams"][:, :2].contiguous())
        got2b = cosa.cam2mask(images=d["img_denorm"], cams=aux_cams, **kw)
    assert torch.equal(got1, want1) and torch.equal(got2, want2) and torch.equal(got3, want3)
    assert torch.equal(got1b, want1) and torch.equal(got2b, want2)
    assert (n2 - n1) == (n1 - n0) - 1, "the second call must skip exactly the affinity launch"
    assert par._shared is None


def test_graphed_step_replays_the_eager_step(cosa):
    """cosa_b200.GraphedStep: the device step captured in a CUDA graph reproduces the eager calls (labels identical,
    loss and gradient equal) for successive batches with different class lists, with one launch per replay."""
    par = cosa.PAR(num_iter=10, dilations=DIL).cuda()
    layer = cosa.DenseEnergyLoss(weight=1e-7, sigma_rgb=15, sigma_xy=100, scale_factor=0.5)
    B, C, H, W = 2, 21, 96, 128
    first = batch(B=B, C=C, H=H, W=W, n_fg=2, seed=61)
    step = cosa.GraphedStep(par, layer, 0.7, 0.25, B=B, C=C, H=H, W=W, img_box=first["img_box"])
    for seed, n_fg in ((61, 2), (62, 3), (63, 1)):
        hb = batch(B=B, C=C, H=H, W=W, n_fg=n_fg, seed=seed)
        d = to_cuda(hb)
        step.simg.copy_(d["simg"]); step.cams.copy_(d["cams"]); step.cls_label.copy_(d["cls_label"])
        with torch.no_grad():
            step.logits.copy_(d["logits"])
        label, loss, grad = step()
        cams = cosa.cam_validation(d["cams"], d["cls_label"])
        want = cosa.cam2mask(images=cosa.denormalize_img(d["simg"]), img_boxes=hb["img_box"], cams=cams,
                             cls_labels=d["cls_label"], threshold_high=0.7, threshold_low=0.25, refine_model=par)
        logit = d["logits"].clone().requires_grad_(True)
        wl = cosa.get_energy_loss(img=d["simg"], logit=logit, label=want, img_box=hb["img_box"], loss_layer=layer)
        wl.backward()
        assert torch.equal(label, want), seed
        assert_close(loss, wl, "graphed loss", tol=1e-5)
        assert_close(grad, logit.grad, "graphed gradient", tol=1e-5)


def test_prebuilt_lattice_matches_the_single_stream_call(cosa):
    """DenseEnergyLoss.prebuild_lattice: the image-only half of get_energy_loss (de-normalise, nearest 2:1, lattice
    build) started on a second stream before cam2mask.  Same vertex count, loss and gradient as the plain call; a
    prebuilt lattice is used once and only by the call it was made for."""
    from cosa_b200 import _lib, seg_helper
    par = cosa.PAR(num_iter=10, dilations=DIL).cuda()
    layer = cosa.DenseEnergyLoss(weight=1e-7, sigma_rgb=15, sigma_xy=100, scale_factor=0.5)
    B, C, H, W = 3, 21, 96, 128
    d = to_cuda(batch(B=B, C=C, H=H, W=W, n_fg=2, seed=71))
    other = to_cuda(batch(B=B, C=C, H=H, W=W, n_fg=2, seed=72))

    def run(prebuild, simg_for_loss):
        if prebuild is not None:
            assert layer.prebuild_lattice(prebuild, C)
        cams = cosa.cam_validation(d["cams"], d["cls_label"])
        label = cosa.cam2mask(images=cosa.denormalize_img(d["simg"]), img_boxes=d["img_box"], cams=cams,
                              cls_labels=d["cls_label"], threshold_high=0.7, threshold_low=0.25, refine_model=par)
        logit = d["logits"].clone().requires_grad_(True)
        n0 = _lib.launch_count()
        loss = cosa.get_energy_loss(img=simg_for_loss, logit=logit, label=label, img_box=d["img_box"], loss_layer=layer)
        n1 = _lib.launch_count()
        loss.backward()
        M = seg_helper.last_energy_lattice_stats(B, C, H, W)[0]
        return label, loss.detach().clone(), logit.grad.clone(), M, n1 - n0

    label0, loss0, grad0, M0, launches0 = run(None, d["simg"])
    label1, loss1, grad1, M1, launches1 = run(d["simg"], d["simg"])
    assert torch.equal(label0, label1) and M0 == M1 and M0 > 0
    assert launches1 == launches0 - 5, "the forward must skip exactly the five build launches"
    assert_close(loss1, loss0, "prebuilt loss", tol=1e-5)
    assert_close(grad1, grad0, "prebuilt gradient", tol=1e-5)
    # a lattice prebuilt for another image tensor is not picked up ...
    _, loss2, grad2, M2, launches2 = run(other["simg"], d["simg"])
    assert launches2 == launches0 and M2 == M0
    assert_close(loss2, loss0, "loss after a mismatched prebuild", tol=1e-5)
    # ... and is dropped: the next plain call builds its own
    _, loss3, _, _, launches3 = run(None, d["simg"])
    assert launches3 == launches0
    assert_close(loss3, loss0, "loss after the dropped prebuild", tol=1e-5)
    # geometries the fused path does not take are refused quietly
    assert layer.prebuild_lattice(d["simg"].cpu(), C) is False
    assert cosa.DenseEnergyLoss(1e-7, 15, 100, 1.0).prebuild_lattice(d["simg"], C) is False


def test_par_two_streams_two_dilation_sets(cosa, port):
    """Two PAR instances with different dilation lists running concurrently on two streams: the dilations and the
    position term travel with each launch (kernel arguments), so neither call can see the other's constants."""
    gen = torch.Generator().manual_seed(77)
    imgs = torch.rand((2, 3, 64, 96), generator=gen)
    masks = torch.rand((2, 4, 64, 96), generator=gen).softmax(dim=1)
    sets = ([1, 2, 4, 8, 12, 24], [3, 5, 7])
    want = [port.par_forward(imgs, masks, tuple(d), 6) for d in sets]
    pars = [cosa.PAR(num_iter=6, dilations=d).cuda() for d in sets]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    di, dm = imgs.cuda(), masks.cuda()
    torch.cuda.synchronize()
    outs = [[], []]
    for _ in range(5):
        for k in (0, 1):
            with torch.cuda.stream(streams[k]):
                outs[k].append(pars[k](di, dm))
    torch.cuda.synchronize()
    for k in (0, 1):
        for o in outs[k]:
            assert_close(o, want[k], "PAR dilations %s on its own stream" % sets[k])


def test_results_do_not_depend_on_stale_scratch(cosa, port):
    """Every entry point must initialise what it reads: the cached scratch buffers are filled with 0xFF bytes (NaN
    floats, -1 indices, all-ones keys) between two runs of the whole path and of the dense-CRF inference, and the
    results must be the ones of the first run."""
    from cosa_b200 import _lib
    host = batch(B=3, C=21, H=90, W=122, n_fg=2, seed=211)        # h*w % 4 != 0 at half resolution: padding keys
    d = to_cuda(host)
    par = cosa.PAR(num_iter=10, dilations=DIL).cuda()
    layer = cosa.DenseEnergyLoss(weight=1e-7, sigma_rgb=15, sigma_xy=100, scale_factor=0.5)
    probs = d["logits"].softmax(dim=1)
    img255 = (d["img_denorm"] * 255).contiguous()

    def run():
        label = cosa.cam2mask(images=cosa.denormalize_img(d["simg"]), img_boxes=host["img_box"],
                              cams=cosa.cam_validation(d["cams"], d["cls_label"]), cls_labels=d["cls_label"],
                              threshold_high=0.7, threshold_low=0.25, refine_model=par)
        logit = d["logits"].clone().requires_grad_(True)
        loss = cosa.get_energy_loss(img=d["simg"], logit=logit, label=label, img_box=host["img_box"], loss_layer=layer)
        loss.backward()
        q = cosa.crf_inference_batch(img255, probs, 2, 1, 1, 4, 121, 5)
        torch.cuda.synchronize()
        return label.clone(), loss.detach().clone(), logit.grad.clone(), q.clone()

    first = run()
    for buf in list(_lib._scratch.values()):
        buf.fill_(255)
    torch.cuda.synchronize()
    second = run()
    assert torch.equal(first[0], second[0])
    assert_close(second[1], first[1], "loss after poisoning the scratch", tol=1e-6)
    assert_close(second[2], first[2], "gradient after poisoning the scratch", tol=1e-5)
    assert_close(second[3], first[3], "dense-CRF marginals after poisoning the scratch", tol=1e-5)


def test_vertex_budget_shrinks_the_scratch_and_overflow_is_loud(cosa, port):
    """DenseEnergyLoss.vertex_budget: the same loss and gradient from a workspace sized for 1.5 vertices per pixel
    instead of the worst case 6; a budget the batch outgrows yields NaN and error flag 2, not an overrun."""
    from cosa_b200 import _lib, seg_helper
    lib = _lib.load()
    B, C, H, W = 4, 21, 128, 160
    d = to_cuda(batch(B=B, C=C, H=H, W=W, n_fg=2, seed=401))
    label = torch.zeros((B, H, W), device="cuda")
    full = lib.cosa_energy_loss_ws_bytes(B, C, H, W)
    small = lib.cosa_energy_loss_ws_bytes_ex(B, C, H, W, (int(1.5 * 16 + 0.5) & 0xff) << 8)
    assert small < 0.45 * full, (small, full)

    def run(budget, prebuild):
        layer = cosa.DenseEnergyLoss(weight=1e-7, sigma_rgb=15, sigma_xy=100, scale_factor=0.5)
        layer.vertex_budget = budget
        if prebuild:
            assert layer.prebuild_lattice(d["simg"], C)
        logit = d["logits"].clone().requires_grad_(True)
        loss = cosa.get_energy_loss(img=d["simg"], logit=logit, label=label, img_box=d["img_box"], loss_layer=layer)
        loss.backward()
        return loss.detach().clone(), logit.grad.clone(), seg_helper.last_energy_lattice_stats(B, C, H, W)

    l0, g0, st0 = run(None, False)
    for pre in (False, True):
        l1, g1, st1 = run(1.5, pre)
        assert st1[0] == st0[0] and st1[1] == 0
        assert_close(l1, l0, "loss with a vertex budget", tol=1e-6)
        assert_close(g1, g0, "gradient with a vertex budget", tol=1e-5)
    l2, g2, st2 = run(0.0625, False)                  # 1/16 vertex per pixel: the batch needs ~0.5
    assert bool(torch.isnan(l2).all()) and st2[1] & 2, (l2, st2)
    l3, _, st3 = run(None, False)                     # and the next plain call is unaffected
    assert st3[1] == 0
    assert_close(l3, l0, "loss after an overflowed call", tol=1e-6)
    with pytest.raises(ValueError):
        run(100.0, False)


def test_cam2mask_class_budget(cosa):
    """cam2mask(max_classes=n): same labels from mask buffers sized for n present classes per image; an image with
    more classes than the budget poisons the call's labels instead of overrunning the scratch."""
    from cosa_b200 import _lib
    lib = _lib.load()
    assert lib.cosa_cam2mask_ws_bytes_ex(32, 20, 448, 448, 2, 1, 6, 6 << 8) < 0.5 * lib.cosa_cam2mask_ws_bytes(32, 20, 448, 448, 2, 1, 6)
    host = batch(B=3, C=21, H=96, W=128, n_fg=3, seed=77)
    d = to_cuda(host)
    par = cosa.PAR(num_iter=10, dilations=DIL).cuda()
    kw = dict(images=d["img_denorm"], img_boxes=host["img_box"], cams=d["cams"], cls_labels=d["cls_label"],
              threshold_high=0.7, threshold_low=0.25)
    for refine in (par, None):
        want = cosa.cam2mask(refine_model=refine, **kw)
        assert torch.equal(cosa.cam2mask(refine_model=refine, max_classes=3, **kw), want)
        assert torch.equal(cosa.cam2mask(refine_model=refine, max_classes=7, **kw), want)
        over = cosa.cam2mask(refine_model=refine, max_classes=2, **kw)          # three classes present, budget two
        assert bool(torch.isnan(over).all())
        assert torch.equal(cosa.cam2mask(refine_model=refine, **kw), want)       # the next call is unaffected


@pytest.mark.parametrize("C", [21, 81, 5])
def test_energy_loss_on_unaligned_tensor_views(cosa, C):
    """Tensor views that start at an odd element (4-byte aligned only): the vectorised energy kernels read 8 / 16 bytes
    at a time, so the dispatch must fall back to the scalar kernels - same loss and gradient as on aligned copies."""
    d = to_cuda(batch(B=2, C=C, H=64, W=96, n_fg=2, seed=311))
    label = cosa.cam_to_label(d["cams"], d["cls_label"], img_box=d["img_box"], bkg_thre=0.35, high_thre=0.55,
                              low_thre=0.35, ignore_mid=True, ignore_index=255)[1].float()

    def off1(t):                      # the same values, one float past a 256-byte aligned allocation
        buf = torch.empty(t.numel() + 1, dtype=t.dtype, device=t.device)
        v = buf[1:].view(t.shape)
        v.copy_(t)
        assert v.is_contiguous() and v.data_ptr() % 8 == 4
        return v

    def run(simg, logits, lab):
        layer = cosa.DenseEnergyLoss(weight=1e-7, sigma_rgb=15, sigma_xy=100, scale_factor=0.5)
        logit = logits.detach().requires_grad_(True)
        loss = cosa.get_energy_loss(img=simg, logit=logit, label=lab, img_box=d["img_box"], loss_layer=layer)
        loss.backward()
        return loss.detach(), logit.grad

    loss0, grad0 = run(d["simg"], d["logits"], label)
    loss1, grad1 = run(off1(d["simg"]), off1(d["logits"]), off1(label))
    assert_close(loss1, loss0, "loss on unaligned views", tol=1e-5)
    assert_close(grad1, grad0, "gradient on unaligned views", tol=1e-5)
