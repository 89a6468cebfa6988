#!/usr/bin/env python
"""BASELINE.json configs[4]: the hot path INSIDE a full co-training step, batch-sharded DDP at 1/2/4/8 B200.

The networks of CoSA (ViT-B/16 student + EMA teacher, timm) are outside the scope of this repository (SURVEY.md 8:
one data-parallel hot path), so they are stood in for by a plain-torch ViT-B/16 of the same shape (12 layers, 768
wide, 12 heads, patch 16, bf16 autocast) with random weights: a classification head on the patch tokens for the CAMs
and a 1x1 decoder for the segmentation logits.  What this script measures is what configs[4] asks about the PATH:

  * images/s of one whole step (main.py:106-252): teacher forward at the three pseudo-label scales with flips
    (multi_scale_camseg, seg_helper.py:232-275) -> merge/normalise -> cam_validation -> cam2mask(+PAR) -> student
    forward -> seg_loss + get_energy_loss (+ classification loss) -> backward (DDP bucketed all-reduce over NCCL)
    -> AdamW step -> EMA update (main.py:49-50, 245-252);
  * the share of that step spent in this repository's kernels (CUDA events around the path calls), and
  * weak scaling over the GPUs of one box (the path has no collective; DDP's gradient all-reduce is the only traffic).

    python bench_full_step.py [--batch 16] [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_full_step.py ...

Rank 0 prints one JSON line.  This is a side measurement: the contract line of the round is bench.py's.
"""
import argparse
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


class ViTStandIn(nn.Module):
    """ViT-B/16-shaped encoder with a CAM head and a segmentation decoder (the interface of the reference's
    VITNetwork.forward(x, cam_only=False) -> (cls, cls_aux, feature_map, seg, cam, cam_aux), main.py:124)."""

    def __init__(self, num_classes=21, dim=768, depth=12, heads=12, patch=16, grid=28):
        super().__init__()
        self.patch, self.grid = patch, grid
        self.embed = nn.Conv2d(3, dim, patch, patch)
        self.pos = nn.Parameter(torch.zeros(1, dim, grid, grid))
        layer = nn.TransformerEncoderLayer(dim, heads, 4 * dim, dropout=0.0, activation="gelu", batch_first=True,
                                           norm_first=True)
        self.blocks = nn.TransformerEncoder(layer, depth, enable_nested_tensor=False)
        self.norm = nn.LayerNorm(dim)
        self.cam_head = nn.Conv2d(dim, num_classes - 1, 1, bias=False)
        self.aux_head = nn.Conv2d(dim, num_classes - 1, 1, bias=False)
        self.decoder = nn.Conv2d(dim, num_classes, 1)
        nn.init.trunc_normal_(self.pos, std=0.02)

    def forward(self, x, cam_only=False):
        t = self.embed(x)
        b, d, h, w = t.shape
        pos = self.pos if (h, w) == self.pos.shape[2:] else F.interpolate(self.pos, size=(h, w), mode="bilinear",
                                                                          align_corners=False)
        t = (t + pos).flatten(2).transpose(1, 2)
        t = self.norm(self.blocks(t)).transpose(1, 2).reshape(b, d, h, w)
        cam, cam_aux = self.cam_head(t), self.aux_head(t)
        cls = F.adaptive_max_pool2d(cam, 1).flatten(1)
        cls_aux = F.adaptive_max_pool2d(cam_aux, 1).flatten(1)
        return cls, cls_aux, t, self.decoder(t), cam, cam_aux


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16, help="images per GPU")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=448)
    ap.add_argument("--classes", type=int, default=21)
    args = ap.parse_args()

    import torch.distributed as dist
    import cosa_b200
    from cosa_b200 import sharding, synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench_full_step.py needs a CUDA device")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cosa_b200._lib.load()
    torch.manual_seed(1234)          # the same initial weights on every rank
    B, C, S = args.batch, args.classes, args.size
    student = ViTStandIn(C, grid=S // 16).to(dev)
    teacher = ViTStandIn(C, grid=S // 16).to(dev)
    teacher.load_state_dict(student.state_dict())
    for p in teacher.parameters():
        p.requires_grad_(False)
    model = nn.parallel.DistributedDataParallel(student, device_ids=[local_rank]) if world > 1 else student
    opt = torch.optim.AdamW(student.parameters(), lr=6e-5, weight_decay=0.01)
    par = cosa_b200.PAR(num_iter=10, dilations=[1, 2, 4, 8, 12, 24]).to(dev)
    layer = cosa_b200.DenseEnergyLoss(weight=1e-7, sigma_rgb=15, sigma_xy=100, scale_factor=0.5)
    host = synthetic.synthetic_batch(B=B, C=C, H=S, W=S, n_fg=2, seed=5000 + rank)
    simg, cls_label, boxes = host["simg"].to(dev), host["cls_label"].to(dev), host["img_box"]
    scales = [1.0, 0.5, 1.5]                                    # pseudo_scales of the reference's VOC config

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    path_ms = []

    def step(timed):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            raw_cam, raw_seg = [], []
            for s in scales:
                x = simg if s == 1.0 else F.interpolate(simg, size=(int(s * S), int(s * S)), mode="bilinear",
                                                        align_corners=False)
                _, _, _, seg_t, cam_t, _ = teacher(torch.cat([x, x.flip(-1)], 0))
                raw_cam.append(cam_t.float())
                raw_seg.append(seg_t.float())
        ev[0].record()
        with torch.no_grad():                                   # ---- the path, part 1: pseudo labels ----
            early = cosa_b200.par.lattice_prebuild_before_cam2mask()
            if early:
                layer.prebuild_lattice(simg, C)
            cams = cosa_b200.multi_scale_cam_merge(raw_cam, (S, S), cls_label=cls_label)
            label = cosa_b200.cam2mask(images=cosa_b200.denormalize_img(simg), img_boxes=boxes, cams=cams,
                                       cls_labels=cls_label, threshold_high=0.7, threshold_low=0.25, refine_model=par)
            if not early:
                layer.prebuild_lattice(simg, C)       # second stream: runs under the student's forward pass
        ev[1].record()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            cls, cls_aux, _, seg, _, _ = model(simg)
        cls_loss = F.multilabel_soft_margin_loss(cls.float(), cls_label) + F.multilabel_soft_margin_loss(cls_aux.float(), cls_label)
        ev[2].record()
        seg_pred = cosa_b200.upsample_bilinear(seg.float(), (S, S))          # ---- the path, part 2: losses ----
        seg_loss = cosa_b200.seg_loss(seg_pred, label, fg_alpha=0.5)
        reg_loss = cosa_b200.get_energy_loss(img=simg, logit=seg_pred, label=label, img_box=boxes, loss_layer=layer)
        ev[3].record()
        loss = cls_loss + 0.1 * seg_loss + 0.05 * reg_loss.sum()
        opt.zero_grad(set_to_none=True)
        loss.backward()                                         # (the path's backward kernels run inside this call)
        opt.step()
        with torch.no_grad():                                   # EMA teacher (main.py:250-252)
            torch._foreach_mul_(list(teacher.parameters()), 0.999)
            torch._foreach_add_(list(teacher.parameters()), list(student.parameters()), alpha=0.001)
        ev[4].record()
        if timed:
            torch.cuda.synchronize()
            path_ms.append(ev[0].elapsed_time(ev[1]) + ev[2].elapsed_time(ev[3]))
        return loss

    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    sharding.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        loss = step(False)
    t1.record()
    torch.cuda.synchronize()
    sharding.barrier()
    ms = sharding.all_reduce_max(t0.elapsed_time(t1)) / args.steps
    for _ in range(3):                      # a few more steps with a sync each, to split the step by CUDA events
        step(True)
    fwd_path = sum(path_ms) / len(path_ms)
    # the path's backward kernels (energy gradient, seg_loss gradient, enlargement adjoint) timed on their own
    seg = torch.randn((B, C, S // 16, S // 16), device=dev, requires_grad=True)
    label = torch.zeros((B, S, S), device=dev)
    sp = cosa_b200.upsample_bilinear(seg, (S, S))
    l = 0.1 * cosa_b200.seg_loss(sp, label) + 0.05 * cosa_b200.get_energy_loss(img=simg, logit=sp, label=label, img_box=boxes,
                                                                                loss_layer=layer).sum()
    torch.cuda.synchronize()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    l.backward()
    b1.record()
    torch.cuda.synchronize()
    bwd_path = b0.elapsed_time(b1)
    total = sharding.all_reduce_sum(B)
    if rank == 0:
        n_params = sum(p.numel() for p in student.parameters())
        print(json.dumps({
            "metric": "full co-training step images/sec (ViT-B/16 stand-in + the refinement path, DDP)",
            "value": total / (ms / 1e3), "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "dtype": "bf16 networks / f32 path",
            "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[4]: %dx%d, %d classes, %d images/GPU x %d GPU; student + EMA "
                                   "teacher = plain-torch ViT-B/16 stand-ins (%.1f M parameters each), teacher at scales "
                                   "%s with flips, AdamW, DDP (NCCL) for N > 1" % (S, S, C, B, world, n_params / 1e6, scales)},
            "path_ms_per_step": {"pseudo_labels_and_losses_forward": fwd_path, "losses_backward": bwd_path},
            "path_share_of_step": (fwd_path + bwd_path) / ms,
            "loss": float(loss.detach()),
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
