"""Label flips of cam2mask + PAR against the CPU oracle, for both channel modes (derived last channel / all channels
propagated), with the oracle's top-1/top-2 margin at the flipped pixels.

    python profiles/label_flips.py [B]
"""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import cosa_b200 as cosa
from cosa_b200 import seg_helper, synthetic
from oracle import reference_port as port
from test_gpu_parity import _oracle_margin
DIL=[1,2,4,8,12,24]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for kind in ("noise", "blobs"):
    d = synthetic.synthetic_batch(B, 21, 448, 448, 2, seed=2000)
    cls = d["cls_label"]
    cls[3] = 0; cls[4] = 0; cls[4, 7] = 1; cls[5, :5] = 1
    if kind == "noise":
        cams = torch.rand(d["cams"].shape, generator=torch.Generator().manual_seed(1)) * cls[:, :, None, None]
    else:
        cams = d["cams"]
        cams[5] = torch.rand(cams[5].shape, generator=torch.Generator().manual_seed(1)) * cls[5, :, None, None]
    args = dict(img_boxes=d["img_box"], threshold_high=0.7, threshold_low=0.25)
    want = port.cam2mask(images=d["img_denorm"], cams=cams, cls_labels=cls, refine_model=port.ParOracle(), **args)
    marg = _oracle_margin(port, dict(images=d["img_denorm"], cams=cams, cls_label=cls), port.ParOracle())
    for mode in (True, False):
        seg_helper.cam2mask_propagate_all_channels(mode)
        got = cosa.cam2mask(images=d["img_denorm"].cuda(), cams=cams.cuda(), cls_labels=cls.cuda(),
                            refine_model=cosa.PAR(num_iter=10, dilations=DIL).cuda(), **args).cpu()
        diff = got != want
        per = [int(diff[i].sum()) for i in range(B)]
        print(kind, "all_channels" if mode else "derived", "flips vs oracle:", int(diff.sum()), per,
              "worst margin %.3g" % (float(marg[diff].max()) if diff.any() else 0.0),
              "pixels with margin<1e-6: %d, <1e-5: %d" % (int((marg < 1e-6).sum()), int((marg < 1e-5).sum())))
seg_helper.cam2mask_propagate_all_channels(False)
