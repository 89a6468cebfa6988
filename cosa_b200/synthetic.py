"""Seeded synthetic VOC/COCO-shaped inputs for the hot path.

Plain torch-CPU tensor generation, no kernels: shared by ``bench.py``, the tests and the golden-vector
script so that the CUDA path and the CPU checker always see the same tensors.

  images   u8 = clamp(floor(127 + 100 sin(0.02x + c) cos(0.03y) + sigma N(0,1)), 0, 255)
           simg = (u8 - mean)/std (ImageNet), img_denorm = u8/255
  cams     "blobs": per present class 1-3 Gaussian blobs + 0.05 U(0,1), min-max normalised, x cls_label
           "grid":  U(0,1) on the ViT token grid, bilinearly up-sampled, x cls_label
  logits   3 N(0,1) on the token grid, bilinearly up-sampled (main.py:167)
  img_box  full image, or a cropped variant
"""
import torch
import torch.nn.functional as F

IMAGENET_MEAN = (123.675, 116.28, 103.53)
IMAGENET_STD = (58.395, 57.12, 57.375)


def synthetic_batch(B, C, H, W, n_fg, seed, noise_sigma=10.0, cam_kind="blobs", box="full"):
    g = torch.Generator().manual_seed(seed)
    ys = torch.arange(H, dtype=torch.float32).view(1, H, 1)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, W)
    cc = torch.arange(3, dtype=torch.float32).view(3, 1, 1)
    base = 127 + 100 * torch.sin(0.02 * xs + cc) * torch.cos(0.03 * ys)
    u8 = (base.unsqueeze(0) + noise_sigma * torch.randn((B, 3, H, W), generator=g)).floor().clamp(0, 255)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    simg = (u8 - mean) / std
    img_denorm = u8 / 255.0
    cls_label = torch.zeros((B, C - 1))
    for b in range(B):
        cls_label[b, torch.randperm(C - 1, generator=g)[:n_fg]] = 1
    if cam_kind == "blobs":
        cams = 0.05 * torch.rand((B, C - 1, H, W), generator=g)
        for b in range(B):
            for c in torch.nonzero(cls_label[b])[:, 0].tolist():
                for _ in range(int(torch.randint(1, 4, (1,), generator=g))):
                    cy, cx = (torch.rand(2, generator=g) * torch.tensor([H, W])).tolist()
                    s = float(torch.rand(1, generator=g)) * (60.0 * H / 448) + 30.0 * H / 448
                    cams[b, c] += torch.exp(-((ys[0] - cy) ** 2 + (xs[0] - cx) ** 2) / (2 * s * s))
        cams = cams + (-cams).amax(dim=(2, 3), keepdim=True)
        cams = cams / (cams.amax(dim=(2, 3), keepdim=True) + 1e-5)
    else:
        gh, gw = max(H // 16, 2), max(W // 16, 2)
        cams = F.interpolate(torch.rand((B, C - 1, gh, gw), generator=g), size=(H, W), mode="bilinear",
                             align_corners=False)
    cams = cls_label[:, :, None, None] * cams
    gh, gw = max(H // 16, 2), max(W // 16, 2)
    seg_lowres = 3 * torch.randn((B, C, gh, gw), generator=g)          # decoder logits on the token grid
    logits = F.interpolate(seg_lowres, size=(H, W), mode="bilinear", align_corners=False)      # main.py:167
    if box == "full":
        boxes = torch.tensor([[0, H, 0, W]] * B, dtype=torch.int16)
    else:
        boxes = torch.tensor([[H // 28, H - H // 28, W // 14, W]] * B, dtype=torch.int16)
    return dict(simg=simg.contiguous(), img_denorm=img_denorm.contiguous(), cams=cams.contiguous(),
                cls_label=cls_label, logits=logits.contiguous(), img_box=boxes, seg_lowres=seg_lowres.contiguous())


def synthetic_raw_cams(d, seed, scales=(1.0, 0.5, 1.5)):
    """Teacher CAMs as ``multi_scale_camseg`` receives them from the model (seg_helper.py:246-250): per scale a tensor
    [2B, C-1, hs, ws] on the ViT token grid, the second half for the horizontally flipped images.  Built from the
    batch's full-resolution CAMs (area-averaged onto the grid, a little noise, absent classes at noise level) so that
    the merged result resembles them."""
    g = torch.Generator().manual_seed(seed)
    cams = d["cams"]
    B, C1, H, W = cams.shape
    out = []
    for s in scales:
        hs, ws = max(int(s * H) // 16, 2), max(int(s * W) // 16, 2)
        base = F.adaptive_avg_pool2d(cams, (hs, ws))
        a = base + 0.05 * torch.rand((B, C1, hs, ws), generator=g)
        f = base.flip(-1) + 0.05 * torch.rand((B, C1, hs, ws), generator=g)
        out.append(torch.cat([a, f], dim=0).contiguous())
    return out
