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
path reads crosses PCIe: ``cam_validation`` multiplies each CAM plane by its class label and ``cam2mask`` reads only
the planes of present classes, so planes whose label is 0 are never uploaded - for VOC shapes (2 of 20 classes
present) that is 90 % of the CAM bytes - and the [0,1] image for cam2mask / PAR is derived on the device from the normalised network input
exactly as the reference does (main.py:117, ``denormalize_img``), so the image crosses once.

One step is exactly the device path of main.py:117-212:
denormalize_img -> cam_validation -> cam2mask(refine_model=par) -> get_energy_loss -> backward.
By default each staging slot's step is captured once in a CUDA graph (``GraphedStep`` on the slot's static device
buffers) and replayed with one launch per batch; ``graph=False`` issues the calls one by one.

``submit_native`` takes the same step one stage further upstream on both inputs, to where the tensors are small:
the teacher's raw multi-scale CAM maps on the ViT token grids (what ``multi_scale_camseg`` gets from the model,
seg_helper.py:246-257) and the decoder's logits on the token grid (main.py:167 enlarges them).  The merge / normalise
and the enlargement then run on the device (``multi_scale_cam_merge``, ``upsample_bilinear``) and 93 MB instead of
668 MB cross PCIe per VOC batch of 32.
"""
import copy

import torch

from . import _lib, seg_helper
from . import par as par_mod

_FULL = ("simg", "cls_label", "logits")


class HostPipeline:

    def __init__(self, par, loss_layer, threshold_high, threshold_low, device=None, depth=2, want_grad=False,
                 graph=True):
        if not torch.cuda.is_available():
            raise RuntimeError("cosa_b200.HostPipeline needs a CUDA device: there is no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.par, self.loss_layer = par, loss_layer
        self.thr = (float(threshold_high), float(threshold_low))
        self.depth = int(depth)
        self.want_grad = bool(want_grad)
        # graph=True: the device step of every staging slot is captured once in a CUDA graph (GraphedStep on the slot's
        # own device buffers) and replayed with ONE launch per batch - ~30 launches and their Python/ctypes cost less
        # per batch, which is what limits many ranks on one host.  The first batch of a new shape pays the capture.
        self.use_graph = bool(graph)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.d2h_stream = torch.cuda.Stream(self.device)     # read-backs overlap the kernels of the next batch
        self.slots = []
        self.n_submitted = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.graph_kernels = 0          # kernels of this library launched through graph replays (not in launch_count)
        self._pending = []

    # -- staging ------------------------------------------------------------------------------------------------
    def _slot(self, i, batch, native):
        """Staging slot i for this batch geometry: static device tensors (+ the captured step) and pinned results."""
        while len(self.slots) < self.depth:
            self.slots.append(None)
        names = ("simg", "cls_label", "seg_lowres") if native else _FULL + ("cams",)
        shapes = {k: tuple(batch[k].shape) for k in names}
        if native:
            shapes["raw_cams"] = tuple(tuple(t.shape) for t in batch["raw_cams"])
        s = self.slots[i]
        if s is not None and s["shapes"] == shapes and s["native"] == native:
            return s
        B, _, H, W = shapes["simg"]
        C = shapes["cls_label"][1] + 1
        main = torch.cuda.current_stream(self.device)
        gs = None
        if self.use_graph:
            main.synchronize()        # a slot being replaced may still be in flight
            gs = GraphedStep(self.par, self.loss_layer, self.thr[0], self.thr[1], B=B, C=C, H=H, W=W,
                             img_box=batch["img_box"], device=self.device,
                             raw_cam_shapes=shapes["raw_cams"] if native else None,
                             seg_lowres_shape=shapes["seg_lowres"] if native else None)
            if native:
                dev = {"simg": gs.simg, "cls_label": gs.cls_label, "seg_lowres": gs.seg_lowres, "raw_cams": gs.raw_cams}
            else:
                dev = {"simg": gs.simg, "cls_label": gs.cls_label, "logits": gs.logits, "cams": gs.cams}
        else:
            dev = {k: torch.empty(shapes[k], dtype=torch.float32, device=self.device) for k in names}
            if native:
                dev["raw_cams"] = [torch.empty(sh, dtype=torch.float32, device=self.device) for sh in shapes["raw_cams"]]
        grad_shape = shapes["seg_lowres"] if native else shapes["logits"]
        s = {"shapes": shapes, "native": native, "dev": dev, "graph": gs,
             "ready": torch.cuda.Event(), "consumed": torch.cuda.Event(), "done": torch.cuda.Event(),
             "label": torch.empty((B, H, W), dtype=torch.float32).pin_memory(),
             "loss": torch.empty(1, dtype=torch.float32).pin_memory(),
             "grad": (torch.empty(grad_shape, dtype=torch.float32).pin_memory() if self.want_grad else None)}
        s["consumed"].record(main)
        self.slots[i] = s
        return s

    def _upload(self, s, batch):
        """Host -> device copies of one batch on the copy stream (overlaps the kernels of the previous batch)."""
        dev = s["dev"]
        with torch.cuda.stream(self.copy_stream), torch.no_grad():
            self.copy_stream.wait_event(s["consumed"])       # the step that last read this slot has finished
            for k in (("simg", "cls_label", "seg_lowres") if s["native"] else _FULL):
                dev[k].copy_(batch[k], non_blocking=True)
                self.h2d_bytes += batch[k].numel() * 4
            if s["native"]:
                for dst, src in zip(dev["raw_cams"], batch["raw_cams"]):
                    dst.copy_(src, non_blocking=True)
                    self.h2d_bytes += src.numel() * 4
            else:
                # cam_validation multiplies each plane by its class label and cam2mask reads only the planes of present
                # classes: the others are never uploaded (nor cleared - nothing reads them)
                cams, lab = batch["cams"], batch["cls_label"]
                present = torch.nonzero(lab).tolist()
                if len(present) * 2 >= lab.numel():
                    dev["cams"].copy_(cams, non_blocking=True)
                    self.h2d_bytes += cams.numel() * 4
                else:
                    for b, c in present:                      # one cudaMemcpyAsync per plane
                        dev["cams"][b, c].copy_(cams[b, c], non_blocking=True)
                    self.h2d_bytes += len(present) * cams.shape[2] * cams.shape[3] * 4
            if s["graph"] is not None:
                s["graph"].set_boxes(batch["img_box"])        # the captured kernels read the boxes from device memory
            s["ready"].record(self.copy_stream)

    def _step(self, s, batch, main):
        """The device step of main.py:117-212 on the slot's tensors: one graph replay, or the calls one by one."""
        if s["graph"] is not None:
            main.wait_event(s["done"])        # the previous results of this slot's static outputs have been read back
            self.graph_kernels += s["graph"].kernels_per_replay
            return s["graph"]()
        d, boxes = s["dev"], batch["img_box"]
        H, W = s["shapes"]["simg"][2:]
        early = par_mod.lattice_prebuild_before_cam2mask()
        if early:
            self.loss_layer.prebuild_lattice(d["simg"], d["cls_label"].shape[1] + 1)   # overlaps cam2mask
        img_denorm = seg_helper.denormalize_img(d["simg"])
        if s["native"]:
            # seg_helper.py:250-270 + cam_validation (main.py:137), absent classes' planes zero-filled
            cams = seg_helper.multi_scale_cam_merge(d["raw_cams"], (H, W), cls_label=d["cls_label"])
            leaf = d["seg_lowres"].detach().requires_grad_(True)
            logit = seg_helper.upsample_bilinear(leaf, (H, W))                           # main.py:167
        else:
            cams = seg_helper.cam_validation(d["cams"], d["cls_label"])
            leaf = logit = d["logits"].detach().requires_grad_(True)
        label = seg_helper.cam2mask(images=img_denorm, img_boxes=boxes, cams=cams, cls_labels=d["cls_label"],
                                    threshold_high=self.thr[0], threshold_low=self.thr[1], refine_model=self.par)
        if not early:
            self.loss_layer.prebuild_lattice(d["simg"], d["cls_label"].shape[1] + 1)   # beside get_energy_loss's first kernel
        loss = seg_helper.get_energy_loss(img=d["simg"], logit=logit, label=label, img_box=boxes,
                                          loss_layer=self.loss_layer)
        loss.backward()
        return label, loss.detach(), leaf.grad

    def _submit(self, batch, native):
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            s = self._slot(self.n_submitted % self.depth, batch, native)
            self._upload(s, batch)
            main.wait_event(s["ready"])
            label, loss, grad = self._step(s, batch, main)
            s["consumed"].record(main)
            self._read_back(s, main, label, loss, grad if self.want_grad else None)
        self._pending.append(s)
        self.n_submitted += 1
        if len(self._pending) > self.depth - 1:              # keep at most depth-1 unread results behind us
            return self._collect(self._pending.pop(0))
        return None

    # -- public -------------------------------------------------------------------------------------------------
    def submit(self, batch):
        """Queue one batch: dict with pinned float32 CPU tensors ``simg`` [B,3,H,W] (ImageNet-normalised),
        ``cams`` [B,C-1,H,W], ``cls_label`` [B,C-1], ``logits`` [B,C,H,W] and ``img_box`` (tensor or list, [B,4])."""
        for k in _FULL + ("cams",):
            t = batch[k]
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("HostPipeline.submit: %s must be a contiguous float32 CPU tensor" % k)
        return self._submit(batch, native=False)

    def submit_native(self, batch):
        """Queue one batch given at the resolution the networks produce it: pinned float32 CPU tensors ``simg``
        [B,3,H,W], ``raw_cams`` (list over scales of [2B,C-1,hs,ws]: the teacher's CAMs for the images and their
        flips, seg_helper.py:246-250), ``seg_lowres`` [B,C,h,w] (decoder logits, main.py:167), ``cls_label`` [B,C-1]
        and ``img_box``.  Returns like ``submit``; with ``want_grad`` the gradient is the one of ``seg_lowres``."""
        for t in [batch[k] for k in ("simg", "cls_label", "seg_lowres")] + list(batch["raw_cams"]):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("HostPipeline.submit_native: inputs must be contiguous float32 CPU tensors")
        return self._submit(batch, native=True)

    def _read_back(self, s, main, label, loss, grad):
        """Device -> pinned host copies of a batch's results on their own stream, behind the batch's kernels."""
        s["computed"] = torch.cuda.Event()
        s["computed"].record(main)
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(s["computed"])
            if s["graph"] is None:
                for t in (label, loss, grad):
                    if t is not None:
                        t.record_stream(self.d2h_stream)      # the allocator must not recycle it before the copy ran
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
        step()                      # one cudaGraphLaunch: ~30 kernels, no per-kernel launch cost, no Python in between
        step.label, step.loss, step.grad      # static output tensors, valid until the next call

    The step is ~30 short launches; for small batches (BASELINE.json configs[0], B = 4) and for many ranks sharing one
    host's cores, the host's launch path, not the GPU, sets the pace.  Inputs and outputs are static device tensors (a
    CUDA graph bakes addresses in).  Class labels AND boxes may change between replays: the per-image channel lists
    are built on the device and the boxes live in a static device tensor that ``set_boxes`` rewrites.

    ``raw_cam_shapes`` / ``seg_lowres_shape`` select the native-resolution form (``HostPipeline.submit_native``): the
    inputs are the teacher's raw multi-scale CAMs and the decoder's token-grid logits, the merge / validation and the
    main.py:167 enlargement (with its adjoint) are part of the captured step, and ``grad`` is the gradient of
    ``seg_lowres``.
    """

    def __init__(self, par, loss_layer, threshold_high, threshold_low, B, C, H, W, img_box, device=None,
                 raw_cam_shapes=None, seg_lowres_shape=None):
        if not torch.cuda.is_available():
            raise RuntimeError("cosa_b200.GraphedStep needs a CUDA device: there is no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        dev = self.device
        self.shape = (B, C, H, W)
        self.native = raw_cam_shapes is not None
        self.simg = torch.zeros((B, 3, H, W), dtype=torch.float32, device=dev)
        self.cls_label = torch.zeros((B, C - 1), dtype=torch.float32, device=dev)
        if self.native:
            self.raw_cams = [torch.zeros(tuple(sh), dtype=torch.float32, device=dev) for sh in raw_cam_shapes]
            self.seg_lowres = torch.zeros(tuple(seg_lowres_shape), dtype=torch.float32, device=dev, requires_grad=True)
        else:
            self.cams = torch.zeros((B, C - 1, H, W), dtype=torch.float32, device=dev)
            self.logits = torch.zeros((B, C, H, W), dtype=torch.float32, device=dev, requires_grad=True)
        self.boxes_dev = _lib.resolve_boxes(img_box, B, H, W, dev).clone()
        boxes = _lib.ResolvedBoxes(self.boxes_dev, B, H, W)
        # a private copy of the layer: the graph bakes in the address of the layer-owned lattice workspace
        # (DenseEnergyLoss.prebuild_lattice), which must not be re-allocated by eager calls on the caller's layer
        loss_layer = copy.copy(loss_layer)         # DenseEnergyLoss.__getstate__ leaves the prebuild scratch behind
        self._keep = (boxes, par, loss_layer)      # the graph reads the boxes' device memory on every replay
        thr = (float(threshold_high), float(threshold_low))

        def body():
            # the lattice needs only the image: it is built on a second stream (a forked branch of the graph), under
            # cam2mask next to the per-step PAR kernel, else beside the first kernel of get_energy_loss
            early = par_mod.lattice_prebuild_before_cam2mask()
            if early:
                loss_layer.prebuild_lattice(self.simg, C)
            img_denorm = seg_helper.denormalize_img(self.simg)
            if self.native:
                # seg_helper.py:250-270 + cam_validation (main.py:137), absent classes' planes zero-filled
                cams = seg_helper.multi_scale_cam_merge(self.raw_cams, (H, W), cls_label=self.cls_label)
                leaf = self.seg_lowres
                logit = seg_helper.upsample_bilinear(leaf, (H, W))                           # main.py:167
            else:
                cams = seg_helper.cam_validation(self.cams, self.cls_label)
                leaf = logit = self.logits
            label = seg_helper.cam2mask(images=img_denorm, img_boxes=boxes, cams=cams, cls_labels=self.cls_label,
                                        threshold_high=thr[0], threshold_low=thr[1], refine_model=par)
            if not early:
                loss_layer.prebuild_lattice(self.simg, C)
            loss = seg_helper.get_energy_loss(img=self.simg, logit=logit, label=label, img_box=boxes,
                                              loss_layer=loss_layer)
            grad, = torch.autograd.grad(loss, leaf)
            return label, loss, grad

        with torch.cuda.device(dev):
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):          # first-call work (function attributes, workspaces) outside the capture
                    body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(self.graph):
                self.label, self.loss, self.grad = body()
            self.kernels_per_replay = _lib.launch_count() - n0     # kernels of this library inside one replay

    def set_boxes(self, img_box):
        """Rewrite the static box tensor (stream-ordered copy on the current stream) for the next replays."""
        B, C, H, W = self.shape
        self.boxes_dev.copy_(_lib.resolve_boxes(img_box, B, H, W, torch.device("cpu")), non_blocking=True)

    def __call__(self):
        self.graph.replay()
        return self.label, self.loss, self.grad
