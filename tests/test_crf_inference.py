"""Dense-CRF mean-field inference (utils/seg_helper.py:905-922, 961-996).

Parity status: UNPINNED against pydensecrf (third party, un-vendored, absent here - oracle/crf_oracle.py).  Pinned:
the 2-D and 5-D permutohedral filters against the reference's own Permutohedral class (tests/golden/crf_inference.npz,
made by tests/golden/make_golden.py through oracle/ref_lattice_shim.cpp), and the update rule as restated from the
published algorithm.  CPU part: the C oracle reproduces the golden filter responses bit for bit and the golden
marginals; GPU part: the CUDA product against the oracle and the golden vectors, <= 1e-4 relative, argmax identical
up to numerical ties."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_inf


def test_oracle_filters_and_marginals_match_the_reference_lattice_class():
    from oracle import crf_oracle as co
    g = load_golden("crf_inference")
    img, probs, vals = g["image"], g["probs"], g["vals"]
    H, W = img.shape[:2]
    assert (H * W) % 4 != 0
    for key, feat in (("filt_gauss1", co.gaussian_features(H, W, 1.0)), ("filt_gauss4", co.gaussian_features(H, W, 4.0)),
                      ("filt_bilateral", co.bilateral_features(img, 121, 5))):
        assert np.array_equal(co.oracle_filter(feat, vals), g[key]), key          # bit-exact, D = 2 and D = 5
    for key, args in (("q_infv2", (1, 1, 1, 4, 121, 5)), ("q_inf_t10", (10, 3, 4, 3, 83, 5)), ("q_iter0", (0, 1, 1, 4, 121, 5))):
        q = co.crf_inference(img, probs, *args)
        assert np.array_equal(q, g[key]), key
        assert np.allclose(q.sum(0), 1.0, atol=1e-5)
    # zero iterations: the clipped, re-normalised input probabilities
    want0 = np.clip(probs, 1e-5, 1.0)
    assert rel_inf(g["q_iter0"], want0 / want0.sum(0, keepdims=True)) < 1e-6


def test_host_mirror_keeps_the_reference_names_and_parameters():
    import cosa_b200 as cosa
    v2 = cosa.crf_inference_infv2
    assert isinstance(v2, cosa.DenseCRF)
    assert (v2.iter_max, v2.pos_w, v2.pos_xy_std, v2.bi_w, v2.bi_xy_std, v2.bi_rgb_std) == (1, 1, 1, 4, 121, 5)
    import inspect
    assert list(inspect.signature(cosa.DenseCRF.__init__).parameters)[1:] == [
        "iter_max", "pos_w", "pos_xy_std", "bi_w", "bi_xy_std", "bi_rgb_std"]
    assert list(inspect.signature(cosa.crf_inference_inf).parameters) == ["img", "probs", "t", "scale_factor", "labels"]


def _check_argmax(q, want, what):
    """argmax identical except where the oracle's own top-1/top-2 margin is a numerical tie."""
    diff = q.argmax(0) != want.argmax(0)
    if diff.any():
        top = np.sort(want, axis=0)
        margin = (top[-1] - top[-2])[diff]
        print("%s: %d argmax differences, worst oracle margin %.3g" % (what, int(diff.sum()), float(margin.max())))
        assert float(margin.max()) <= 1e-5


@pytest.mark.gpu
def test_gpu_crf_inference_golden():
    import cosa_b200 as cosa
    g = load_golden("crf_inference")
    img, probs = g["image"], g["probs"]
    q = cosa.crf_inference_infv2(img, probs)                     # the evaluation_engine.py:208 call: numpy in, numpy out
    assert isinstance(q, np.ndarray) and q.shape == probs.shape and q.dtype == np.float32
    r1 = rel_inf(q, g["q_infv2"])
    q10 = cosa.crf_inference_inf(img, probs, t=10, scale_factor=1, labels=probs.shape[0])
    r10 = rel_inf(q10, g["q_inf_t10"])
    q0 = cosa.DenseCRF(0, 1, 1, 4, 121, 5)(img, probs)
    r0 = rel_inf(q0, g["q_iter0"])
    print("dense-CRF marginals vs golden (reference lattice class + restated update): infv2 %.2g, inf t=10 %.2g, "
          "iter_max=0 %.2g" % (r1, r10, r0))
    assert r1 <= 1e-4 and r10 <= 1e-4 and r0 <= 1e-5
    _check_argmax(q, g["q_infv2"], "infv2")
    _check_argmax(q10, g["q_inf_t10"], "inf t=10")
    # CUDA tensors in -> CUDA tensor out, same values
    qd = cosa.crf_inference_infv2(torch.from_numpy(img).cuda(), torch.from_numpy(probs).cuda())
    assert qd.is_cuda and rel_inf(qd.cpu().numpy(), q) <= 1e-6
    with pytest.raises(ValueError):
        cosa.crf_inference_inf(img, probs, labels=probs.shape[0] + 1)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(1, 21, 366, 500, 1), (3, 21, 128, 160, 2), (1, 81, 97, 61, 1)])
def test_gpu_crf_inference_vs_oracle(case):
    """VOC-sized evaluation image (evaluation_engine.py:205-211), a small batch through one lattice pair, and a COCO
    class count on an odd geometry (H*W % 4 != 0, partial tiles): the batched CUDA call against the CPU oracle."""
    import cosa_b200 as cosa
    from oracle import crf_oracle as co
    N, C, H, W, iters = case
    rng = np.random.default_rng(H * W + C)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    imgs = np.stack([np.clip(np.floor(np.stack([127 + 100 * np.sin(0.02 * xx + c + b) * np.cos(0.03 * yy) for c in range(3)], -1)
                                      + 10 * rng.standard_normal((H, W, 3))), 0, 255) for b in range(N)]).astype(np.uint8)
    low = torch.from_numpy(3 * rng.standard_normal((N, C, max(H // 16, 2), max(W // 16, 2))).astype(np.float32))
    probs = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False).softmax(dim=1)
    q = cosa.crf_inference_batch(torch.from_numpy(imgs).permute(0, 3, 1, 2).float().cuda(), probs.cuda(), iters, 1, 1, 4, 121, 5)
    for b in range(N):
        want = co.crf_inference(imgs[b], probs[b].numpy(), iters, 1, 1, 4, 121, 5)
        r = rel_inf(q[b].cpu().numpy(), want)
        print("case %s image %d: marginals rel %.2g" % (case, b, r))
        assert r <= 1e-4
        _check_argmax(q[b].cpu().numpy(), want, "case %s image %d" % (case, b))
