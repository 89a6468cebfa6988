#!/usr/bin/env python
"""Measurements for the rows next to the hot path, same bar as bench.py:

    multi_scale merge (teacher post-processing)   utils/seg_helper.py:253-273
    denormalize_img                               utils/torch_helper.py:354-367 (main.py:117)
    upsample_bilinear forward + adjoint           main.py:167 (F.interpolate of the logits)
    seg_loss forward + backward                   utils/seg_helper.py:800-813
    seg_refine_by_label                           utils/seg_helper.py:553-568
    cam_loss forward + backward                   utils/seg_helper.py:593-602
    dense-CRF mean-field inference (rank 4)       utils/seg_helper.py:961-996 (evaluation_engine.py:205-211)

    python bench_stages.py [--batch 32] [--steps 20] [--warmup 5] [--no-cpu]

One JSON line per stage: device time per call (CUDA events over `steps` calls after `warmup`, inputs resident in
HBM and larger than L2), algorithmic bytes, achieved GB/s against the measured HBM peak, and the CPU oracle port
timed on the host cores on a bounded sample (first 4 images).  VOC shape: 448x448, 21 classes, ViT-B/16 token
grids 28x28 / 14x14 / 42x42 for the scales 1.0 / 0.5 / 1.5.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("bench_stages.py: no CUDA device - cosa_b200 has no CPU fallback")
    import cosa_b200
    from bench import measured_peak_gbs
    from oracle import reference_port as port

    cosa_b200._lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, C, H, W = args.batch, 21, 448, 448
    g = torch.Generator().manual_seed(5)
    raw_cam = [torch.randn((2 * B, C - 1, s, s), generator=g) for s in (28, 14, 42)]
    raw_seg = [torch.randn((2 * B, C, s, s), generator=g) for s in (28, 14, 42)]
    logits = 3 * torch.randn((B, C, H, W), generator=g)
    label = torch.randint(0, C, (B, H, W), generator=g).float()
    label[torch.rand((B, H, W), generator=g) < 0.3] = 255
    label[torch.rand((B, H, W), generator=g) < 0.5] = 0
    cls = torch.zeros((B, C - 1))
    for b in range(B):
        cls[b, torch.randperm(C - 1, generator=g)[:2]] = 1
    cam_pred = torch.randn((B, C - 1, 28, 28), generator=g)
    seg_low = 3 * torch.randn((B, C, 28, 28), generator=g)
    simg = torch.randn((B, 3, H, W), generator=g)
    g_full = torch.randn((B, C, H, W), generator=g).to(dev)
    d = dict(seg_low=seg_low.to(dev), simg=simg.to(dev), raw_cam=[t.to(dev) for t in raw_cam], raw_seg=[t.to(dev) for t in raw_seg], logits=logits.to(dev),
             label=label.to(dev), cls=cls.to(dev), cam_pred=cam_pred.to(dev))
    valid_seg = cosa_b200.seg_refine_by_label(d["logits"], d["cls"], 0.01)
    peak, peak_src = measured_peak_gbs()
    f32 = 4

    def seg_loss_fb(dd, mod):
        x = dd["logits"].detach().requires_grad_(True)
        mod.seg_loss(x, dd["label"], fg_alpha=0.5).backward()
        return x.grad

    def cam_loss_fb(dd, mod, vs):
        x = dd["cam_pred"].detach().requires_grad_(True)
        mod.cam_loss(x, vs).backward()
        return x.grad

    def upsample_fb(x, gfull, mod):
        x = x.detach().requires_grad_(True)
        mod.upsample_bilinear(x, (H, W)).backward(gfull)
        return x.grad

    stages = [
        # name, device fn, cpu fn (on n images), algorithmic bytes per call
        ("denormalize_img", lambda: cosa_b200.denormalize_img(d["simg"], lazy=False), lambda n: port.denormalize_img(simg[:n]),
         f32 * B * 3 * H * W * 2),
        ("upsample_bilinear fwd (28^2 -> 448^2)", lambda: cosa_b200.upsample_bilinear(d["seg_low"], (H, W)), None,
         f32 * B * C * H * W),
        ("upsample_bilinear fwd+bwd (28^2 -> 448^2)", lambda: upsample_fb(d["seg_low"], g_full, cosa_b200),
         lambda n: upsample_fb(seg_low[:n], g_full[:n].cpu(), port),
         f32 * B * C * H * W * 2),                                           # full-size logits written, gradient read
        ("multi_scale_cam_merge", lambda: cosa_b200.multi_scale_cam_merge(d["raw_cam"], (H, W)),
         lambda n: port.multi_scale_merge([t[list(range(n)) + list(range(B, B + n))] for t in raw_cam],
                                          raw_cam[2][list(range(n)) + list(range(B, B + n))],
                                          [t[list(range(n)) + list(range(B, B + n))] for t in raw_seg], (H, W)),
         f32 * B * (C - 1) * H * W + f32 * sum(t.numel() for t in raw_cam)),
        ("multi_scale_seg_merge", lambda: cosa_b200.multi_scale_seg_merge(d["raw_seg"], (H, W)), None,
         f32 * B * C * H * W + f32 * sum(t.numel() for t in raw_seg)),
        ("seg_loss fwd+bwd", lambda: seg_loss_fb(d, cosa_b200),
         lambda n: seg_loss_fb(dict(logits=logits[:n], label=label[:n]), port),
         f32 * B * H * W * (3 * C + 2)),                                   # logits read twice, grad written, label twice
        ("seg_refine_by_label", lambda: cosa_b200.seg_refine_by_label(d["logits"], d["cls"], 0.01),
         lambda n: port.seg_refine_by_label(logits[:n], cls[:n], 0.01), f32 * B * H * W * 2 * C),
        ("cam_loss fwd+bwd", lambda: cam_loss_fb(d, cosa_b200, valid_seg),
         lambda n: cam_loss_fb(dict(cam_pred=cam_pred[:n]), port, valid_seg[:n].cpu()),
         f32 * B * (C - 1) * 28 * 28 * (4 * 1 + 4)),                       # 4 taps + cam, target, grad (token grid)
    ]
    # dense-CRF mean-field inference at evaluation time (seg_helper.py:961-996): one VOC-sized image
    # as evaluation_engine.py:205-211 calls it, and a batch of 16 through one lattice pair.  Algorithmic bytes: the two
    # lattice filters over K = 21 channels (build + norm filter + splat / blur / slice) are data dependent (M); the
    # figure charged here is the floor of the mean-field update itself: probabilities in, marginals out, image in.
    from oracle import crf_oracle
    img_u8 = (torch.rand((16, 3, 448, 448), generator=g) * 255).floor()
    probs16 = logits[:16].softmax(dim=1).to(dev)
    img16 = img_u8.to(dev)
    img1 = img16[:1, :, :366, :].contiguous()
    img1 = torch.nn.functional.pad(img1, (0, 52))[:, :, :, :500].contiguous()
    probs1 = torch.nn.functional.pad(probs16[:1, :, :366, :], (0, 52), value=1.0 / C)[:, :, :, :500].contiguous()
    stages += [
        ("crf_inference_infv2 (1 image 366x500, 21 classes, 1 iteration)",
         lambda: cosa_b200.crf_inference_batch(img1, probs1, 1, 1, 1, 4, 121, 5),
         lambda n: crf_oracle.crf_inference(img1[0].permute(1, 2, 0).cpu().numpy(), probs1[0].cpu().numpy(), 1, 1, 1, 4, 121, 5),
         f32 * 366 * 500 * (2 * C + 3)),
        ("crf_inference_infv2 (16 images 448x448, 21 classes, 1 iteration)",
         lambda: cosa_b200.crf_inference_batch(img16, probs16, 1, 1, 1, 4, 121, 5), None,
         f32 * 16 * 448 * 448 * (2 * C + 3)),
    ]
    for name, fn, cpu_fn, alg in stages:
        for _ in range(max(3, args.warmup)):
            fn()
        torch.cuda.synchronize()
        l0 = cosa_b200._lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        line = {"stage": name, "ms_per_call": round(ms, 4), "images_per_s": round(B / (ms / 1e3), 1),
                "alg_bytes_per_call": alg, "achieved_gbs": round(alg / 1e9 / (ms / 1e3), 1),
                "frac_of_hbm_peak": round(alg / 1e9 / (ms / 1e3) / peak, 4), "peak_gbs": peak, "peak_source": peak_src,
                "gpu_launches_per_call": (cosa_b200._lib.launch_count() - l0) / args.steps,
                "config": "VOC shape B=%d, %dx%d, %d classes" % (B, H, W, C), "cpu": None}
        if cpu_fn is not None and not args.no_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            n = 1 if name.startswith("crf_inference") else min(4, B)
            cpu_fn(1)
            t0 = time.perf_counter()
            cpu_fn(n)
            dt = time.perf_counter() - t0
            line["cpu"] = {"images_per_s": round(n / dt, 2), "cores": os.cpu_count() or 1, "kind": "port",
                           "sample": "first %d images, 1 warm-up image + 1 timed pass" % n}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
