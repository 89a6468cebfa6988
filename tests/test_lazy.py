"""cosa_b200._lazy.LazyTensor: the deferred results of denormalize_img / cam_validation behave as plain tensors
(CPU part: the tensor-subclass mechanics; GPU part: folded into cam2mask bit for bit, materialised by anything else)."""
import pytest
import torch
import torch.nn.functional as F

from cosa_b200._lazy import LazyTensor, pending, plain


def test_lazy_tensor_mechanics_cpu():
    src = torch.arange(24.0).reshape(2, 3, 4)
    calls = []

    def produce():
        calls.append(1)
        return src * 2

    t = LazyTensor("demo", (src,), produce, src)
    assert isinstance(t, torch.Tensor) and t.shape == src.shape and t.dtype == src.dtype and t.device == src.device
    assert pending(t, "demo") == (src,) and pending(t, "other") is None and not calls
    assert "pending" in repr(t)
    assert torch.equal(t + 1, src * 2 + 1) and calls == [1]              # any torch op materialises, once
    assert pending(t, "demo") is None and t.is_materialized
    assert torch.equal(t[1], src[1] * 2) and torch.equal(t.clone(), src * 2) and float(t.sum()) == float(src.sum() * 2)
    assert torch.equal(F.interpolate(t[None], scale_factor=2.0), F.interpolate((src * 2)[None], scale_factor=2.0))
    assert torch.equal(torch.cat([t, t]), torch.cat([src * 2, src * 2])) and calls == [1]
    assert plain(t) is t.materialize() and plain(src) is src
    assert t.clone().numpy().shape == (2, 3, 4)


@pytest.mark.gpu
def test_lazy_producers_fold_into_cam2mask_bit_exact():
    import cosa_b200 as cosa
    from cosa_b200 import synthetic
    host = synthetic.synthetic_batch(B=3, C=21, H=96, W=128, n_fg=2, seed=91)
    host["cls_label"][1] = 0                                   # an image without foreground
    d = {k: (v.cuda() if k != "img_box" else v) for k, v in host.items()}
    raw = d["cams"] + 3.0 * (1 - d["cls_label"])[:, :, None, None]     # garbage in the planes of absent classes
    par = cosa.PAR(num_iter=10, dilations=[1, 2, 4, 8, 12, 24]).cuda()
    kw = dict(img_boxes=host["img_box"], cls_labels=d["cls_label"], threshold_high=0.7, threshold_low=0.25)
    for refine in (par, None):
        for downscale in (2, 0):          # exact-2x kernels and the general ones
            eager = cosa.cam2mask(images=cosa.denormalize_img(d["simg"], lazy=False),
                                  cams=cosa.cam_validation(raw, d["cls_label"], lazy=False), refine_model=refine,
                                  downscale=downscale, **kw)
            li, lc = cosa.denormalize_img(d["simg"]), cosa.cam_validation(raw, d["cls_label"])
            assert isinstance(li, LazyTensor) and isinstance(lc, LazyTensor)
            fused = cosa.cam2mask(images=li, cams=lc, refine_model=refine, downscale=downscale, **kw)
            assert torch.equal(fused, eager), (refine is not None, downscale)
            assert not li.is_materialized and not lc.is_materialized      # nothing was written to HBM for them
    # a label tensor other than the one the CAMs were validated with: no folding, same result as the eager product
    other = d["cls_label"].clone()
    lc = cosa.cam_validation(raw, d["cls_label"])
    got = cosa.cam2mask(images=cosa.denormalize_img(d["simg"]), cams=lc, refine_model=par,
                        **dict(kw, cls_labels=other))
    assert lc.is_materialized
    assert torch.equal(got, cosa.cam2mask(images=cosa.denormalize_img(d["simg"], lazy=False),
                                          cams=cosa.cam_validation(raw, d["cls_label"], lazy=False), refine_model=par, **kw))
    # any other consumer sees the plain tensors of the reference
    li, lc = cosa.denormalize_img(d["simg"]), cosa.cam_validation(raw, d["cls_label"])
    assert torch.equal(li.clone(), cosa.denormalize_img(d["simg"], lazy=False))
    assert torch.equal(F.interpolate(lc, scale_factor=0.5), F.interpolate(cosa.cam_validation(raw, d["cls_label"], lazy=False),
                                                                           scale_factor=0.5))
    assert torch.equal(lc, raw * d["cls_label"][:, :, None, None])
    # a generic refine_model callable takes the per-image path on materialised tensors
    gen = cosa.cam2mask(images=cosa.denormalize_img(d["simg"]), cams=cosa.cam_validation(raw, d["cls_label"]),
                        refine_model=lambda im, cm: cm, **kw)
    assert torch.equal(gen, cosa.cam2mask(images=cosa.denormalize_img(d["simg"], lazy=False),
                                          cams=cosa.cam_validation(raw, d["cls_label"], lazy=False), **kw))
