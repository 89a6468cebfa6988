"""Host-buffer front end of the hot path: batches in pinned host memory in, labels and loss in pinned host memory out.

The reference keeps its tensors on the GPU between the network and these functions (main.py:121-212) but crosses
to the host for the CRF filter every step (seg_helper.py:884-888).  When the producer of the CAMs lives on the
host side of PCIe (a data-loading or serving process handing over buffers), this class is the entry point:

    pipe = HostPipeline(par, loss_layer, threshold_high=0.7, threshold_low=0.25, device="cuda:0")
    for batch in batches:                 # dicts of pinned CPU tensors
        pipe.submit(batch)
    results = pipe.drain()                # [(label [B,H,W] float32, loss [1]) ...] pinned CPU tensors

``submit`` returns the result of the batch queued ``depth - 1`` calls earlier (None at first) and waits only for
that batch; the tensors are the pipeline's own pinned buffers and stay valid until the next ``submit``.

The host->device copies of batch i+1 run on a copy stream while the kernels of batch i run on the compute stream
(``depth`` staging slots), and the labels and the loss of a batch are copied back asynchronously.  Only what the
path reads crosses PCIe: ``cam_validation`` multiplies each CAM plane by its class label, so planes whose label is 0
are never uploaded (their device copy is cleared instead) - for VOC shapes (2 of 20 classes present) that is 90 % of
the CAM bytes - and the [0,1] image for cam2mask / PAR is derived on the device from the normalised network input
exactly as the reference does (main.py:117, ``denormalize_img``), so the image crosses once.

One step is exactly the device path of main.py:117-212:
denormalize_img -> cam_validation -> cam2mask(refine_model=par) -> get_energy_loss -> backward.

``submit_native`` takes the same step one stage further upstream on both inputs, to where the tensors are small:
the teacher's raw multi-scale CAM maps on the ViT token grids (what ``multi_scale_camseg`` gets from the model,
seg_helper.py:246-257) and the decoder's logits on the token grid (main.py:167 enlarges them).  The merge / normalise
and the enlargement then run on the device (``multi_scale_cam_merge``, ``upsample_bilinear``) and 93 MB instead of
668 MB cross PCIe per VOC batch of 32.
"""
import copy

import torch

from . import _lib, seg_helper

_FULL = ("simg", "cls_label", "logits")


class HostPipeline:

    def __init__(self, par, loss_layer, threshold_high, threshold_low, device=None, depth=2, want_grad=False):
        if not torch.cuda.is_available():
            raise RuntimeError("cosa_b200.HostPipeline needs a CUDA device: there is no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.par, self.loss_layer = par, loss_layer
        self.thr = (float(threshold_high), float(threshold_low))
        self.depth = int(depth)
        self.want_grad = bool(want_grad)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.d2h_stream = torch.cuda.Stream(self.device)     # read-backs overlap the kernels of the next batch
        self.slots = []
        self.n_submitted = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._pending = []

    # -- staging ------------------------------------------------------------------------------------------------
    def _slot(self, i, batch):
        while len(self.slots) < self.depth:
            self.slots.append(None)
        s = self.slots[i]
        shapes = {k: tuple(batch[k].shape) for k in _FULL + ("cams",)}
        if s is None or s["shapes"] != shapes:
            dev = {k: torch.empty(shapes[k], dtype=torch.float32, device=self.device) for k in shapes}
            B, _, H, W = shapes["cams"]
            s = {"shapes": shapes, "dev": dev, "ready": torch.cuda.Event(), "consumed": torch.cuda.Event(),
                 "label": torch.empty((B, H, W), dtype=torch.float32).pin_memory(),
                 "loss": torch.empty(1, dtype=torch.float32).pin_memory(),
                 "grad": (torch.empty(shapes["logits"], dtype=torch.float32).pin_memory() if self.want_grad else None),
                 "done": torch.cuda.Event()}
            s["consumed"].record(torch.cuda.current_stream(self.device))
            self.slots[i] = s
        return s

    def _upload(self, s, batch):
        dev = s["dev"]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(s["consumed"])       # the step that last read this slot has finished
            for k in _FULL:
                dev[k].copy_(batch[k], non_blocking=True)
                self.h2d_bytes += batch[k].numel() * 4
            # CAM planes of absent classes are zero after cam_validation: clear them on the device, upload the rest
            cams, lab = batch["cams"], batch["cls_label"]
            present = torch.nonzero(lab).tolist()
            if len(present) * 2 >= lab.numel():
                dev["cams"].copy_(cams, non_blocking=True)
                self.h2d_bytes += cams.numel() * 4
            else:
                dev["cams"].zero_()
                for b, c in present:
                    dev["cams"][b, c].copy_(cams[b, c], non_blocking=True)
                self.h2d_bytes += len(present) * cams.shape[2] * cams.shape[3] * 4
            s["ready"].record(self.copy_stream)

    # -- public -------------------------------------------------------------------------------------------------
    def submit(self, batch):
        """Queue one batch: dict with pinned float32 CPU tensors ``simg`` [B,3,H,W] (ImageNet-normalised),
        ``cams`` [B,C-1,H,W], ``cls_label`` [B,C-1], ``logits`` [B,C,H,W] and ``img_box`` (tensor or list, [B,4])."""
        for k in _FULL + ("cams",):
            t = batch[k]
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("HostPipeline.submit: %s must be a contiguous float32 CPU tensor" % k)
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            s = self._slot(self.n_submitted % self.depth, batch)
            self._upload(s, batch)
            main.wait_event(s["ready"])
            d, boxes = s["dev"], batch["img_box"]
            self.loss_layer.prebuild_lattice(d["simg"], d["cls_label"].shape[1] + 1)   # overlaps cam2mask
            img_denorm = seg_helper.denormalize_img(d["simg"])
            cams = seg_helper.cam_validation(d["cams"], d["cls_label"])
            label = seg_helper.cam2mask(images=img_denorm, img_boxes=boxes, cams=cams, cls_labels=d["cls_label"],
                                        threshold_high=self.thr[0], threshold_low=self.thr[1], refine_model=self.par)
            logit = d["logits"].detach().requires_grad_(True)
            loss = seg_helper.get_energy_loss(img=d["simg"], logit=logit, label=label, img_box=boxes,
                                              loss_layer=self.loss_layer)
            loss.backward()
            s["consumed"].record(main)
            self._read_back(s, main, label, loss.detach(), logit.grad if self.want_grad else None)
        self._pending.append(s)
        self.n_submitted += 1
        if len(self._pending) > self.depth - 1:              # keep at most depth-1 unread results behind us
            return self._collect(self._pending.pop(0))
        return None

    # -- native-resolution inputs ---------------------------------------------------------------------------------
    def submit_native(self, batch):
        """Queue one batch given at the resolution the networks produce it: pinned float32 CPU tensors ``simg``
        [B,3,H,W], ``raw_cams`` (list over scales of [2B,C-1,hs,ws]: the teacher's CAMs for the images and their
        flips, seg_helper.py:246-250), ``seg_lowres`` [B,C,h,w] (decoder logits, main.py:167), ``cls_label`` [B,C-1]
        and ``img_box``.  Returns like ``submit``; with ``want_grad`` the gradient is the one of ``seg_lowres``."""
        names = ("simg", "cls_label", "seg_lowres")
        for t in [batch[k] for k in names] + list(batch["raw_cams"]):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("HostPipeline.submit_native: inputs must be contiguous float32 CPU tensors")
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            i = self.n_submitted % self.depth
            while len(self.slots) < self.depth:
                self.slots.append(None)
            shapes = {k: tuple(batch[k].shape) for k in names}
            shapes["raw_cams"] = tuple(tuple(t.shape) for t in batch["raw_cams"])
            s = self.slots[i]
            if s is None or s["shapes"] != shapes:
                B, _, H, W = shapes["simg"]
                dev = {k: torch.empty(shapes[k], dtype=torch.float32, device=self.device) for k in names}
                dev["raw_cams"] = [torch.empty(sh, dtype=torch.float32, device=self.device) for sh in shapes["raw_cams"]]
                s = {"shapes": shapes, "dev": dev, "ready": torch.cuda.Event(), "consumed": torch.cuda.Event(),
                     "label": torch.empty((B, H, W), dtype=torch.float32).pin_memory(),
                     "loss": torch.empty(1, dtype=torch.float32).pin_memory(),
                     "grad": (torch.empty(shapes["seg_lowres"], dtype=torch.float32).pin_memory()
                              if self.want_grad else None),
                     "done": torch.cuda.Event()}
                s["consumed"].record(main)
                self.slots[i] = s
            d = s["dev"]
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(s["consumed"])
                for k in names:
                    d[k].copy_(batch[k], non_blocking=True)
                    self.h2d_bytes += batch[k].numel() * 4
                for dst, src in zip(d["raw_cams"], batch["raw_cams"]):
                    dst.copy_(src, non_blocking=True)
                    self.h2d_bytes += src.numel() * 4
                s["ready"].record(self.copy_stream)
            main.wait_event(s["ready"])
            boxes = batch["img_box"]
            H, W = shapes["simg"][2:]
            self.loss_layer.prebuild_lattice(d["simg"], d["cls_label"].shape[1] + 1)   # overlaps cam2mask
            img_denorm = seg_helper.denormalize_img(d["simg"])
            # seg_helper.py:250-270 + cam_validation (main.py:137), absent classes' planes zero-filled
            cams = seg_helper.multi_scale_cam_merge(d["raw_cams"], (H, W), cls_label=d["cls_label"])
            label = seg_helper.cam2mask(images=img_denorm, img_boxes=boxes, cams=cams, cls_labels=d["cls_label"],
                                        threshold_high=self.thr[0], threshold_low=self.thr[1], refine_model=self.par)
            low = d["seg_lowres"].detach().requires_grad_(True)
            logit = seg_helper.upsample_bilinear(low, (H, W))                           # main.py:167
            loss = seg_helper.get_energy_loss(img=d["simg"], logit=logit, label=label, img_box=boxes,
                                              loss_layer=self.loss_layer)
            loss.backward()
            s["consumed"].record(main)
            self._read_back(s, main, label, loss.detach(), low.grad if self.want_grad else None)
        self._pending.append(s)
        self.n_submitted += 1
        if len(self._pending) > self.depth - 1:
            return self._collect(self._pending.pop(0))
        return None

    def _read_back(self, s, main, label, loss, grad):
        """Device -> pinned host copies of a batch's results on their own stream, behind the batch's kernels."""
        s["computed"] = torch.cuda.Event()
        s["computed"].record(main)
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(s["computed"])
            for t in (label, loss, grad):
                if t is not None:
                    t.record_stream(self.d2h_stream)          # the allocator must not recycle it before the copy ran
            s["label"].copy_(label, non_blocking=True)
            s["loss"].copy_(loss, non_blocking=True)
            self.d2h_bytes += s["label"].numel() * 4 + 4
            if grad is not None:
                s["grad"].copy_(grad, non_blocking=True)
                self.d2h_bytes += s["grad"].numel() * 4
            s["done"].record(self.d2h_stream)

    def _collect(self, s):
        s["done"].synchronize()
        out = (s["label"], s["loss"])                        # the slot's pinned buffers: valid until the next submit
        return out + ((s["grad"],) if self.want_grad else ())

    def drain(self):
        """Wait for every queued batch; returns the results not yet handed out by ``submit`` (oldest first)."""
        out = [self._collect(s) for s in self._pending]
        self._pending = []
        return out


class GraphedStep:
    """The device step of main.py:117-212 captured ONCE in a CUDA graph and replayed per batch.

        step = GraphedStep(par, loss_layer, 0.7, 0.25, B=4, C=21, H=448, W=448, img_box=boxes)
        step.simg.copy_(...); step.cams.copy_(...); step.cls_label.copy_(...); step.logits.copy_(...)
        step()                      # one cudaGraphLaunch: ~35 kernels, no per-kernel launch cost, no Python in between
        step.label, step.loss, step.grad      # static output tensors, valid until the next call

    The step is 34 short launches; for small batches (BASELINE.json configs[0], B = 4: 0.55 ms per step) the host's
    launch path, not the GPU, sets the pace.  Inputs and outputs are static device tensors (a CUDA graph bakes
    addresses in); the boxes are resolved once (they are part of the captured launch parameters).  Class labels may
    change between replays: the per-image channel lists are built on the device.
    """

    def __init__(self, par, loss_layer, threshold_high, threshold_low, B, C, H, W, img_box, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("cosa_b200.GraphedStep needs a CUDA device: there is no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        dev = self.device
        self.simg = torch.zeros((B, 3, H, W), dtype=torch.float32, device=dev)
        self.cams = torch.zeros((B, C - 1, H, W), dtype=torch.float32, device=dev)
        self.cls_label = torch.zeros((B, C - 1), dtype=torch.float32, device=dev)
        self.logits = torch.zeros((B, C, H, W), dtype=torch.float32, device=dev, requires_grad=True)
        boxes = _lib.ResolvedBoxes(_lib.resolve_boxes(img_box, B, H, W, dev), B, H, W)
        # a private copy of the layer: the graph bakes in the address of the layer-owned lattice workspace
        # (DenseEnergyLoss.prebuild_lattice), which must not be re-allocated by eager calls on the caller's layer
        loss_layer = copy.copy(loss_layer)         # DenseEnergyLoss.__getstate__ leaves the prebuild scratch behind
        self._keep = (boxes, par, loss_layer)      # the graph reads the boxes' device memory on every replay
        thr = (float(threshold_high), float(threshold_low))

        def body():
            # the lattice needs only the image: built on a second stream (a forked branch of the graph) under cam2mask
            loss_layer.prebuild_lattice(self.simg, C)
            img_denorm = seg_helper.denormalize_img(self.simg)
            cams = seg_helper.cam_validation(self.cams, self.cls_label)
            label = seg_helper.cam2mask(images=img_denorm, img_boxes=boxes, cams=cams, cls_labels=self.cls_label,
                                        threshold_high=thr[0], threshold_low=thr[1], refine_model=par)
            loss = seg_helper.get_energy_loss(img=self.simg, logit=self.logits, label=label, img_box=boxes,
                                              loss_layer=loss_layer)
            grad, = torch.autograd.grad(loss, self.logits)
            return label, loss, grad

        with torch.cuda.device(dev):
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):          # first-call work (function attributes, constants, workspaces) outside the capture
                    body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.label, self.loss, self.grad = body()
        self.kernels_per_replay = None

    def __call__(self):
        self.graph.replay()
        return self.label, self.loss, self.grad
