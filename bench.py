#!/usr/bin/env python
"""Headline benchmark: refined pseudo-label images/sec through PAR + labelling + dense-CRF loss.

One "step" is one pass of the hot path over one synthetic VOC-shaped batch per GPU
(BASELINE.json configs[1]: B=32, 3x448x448 images, 21 classes, 2 foreground classes per image):

    cam_validation -> cam2mask(refine_model=PAR(num_iter=10, dilations=[1,2,4,8,12,24]))
                   -> get_energy_loss(DenseEnergyLoss(1e-7, 15, 100, 0.5)) forward -> backward (d loss/d logit)

    python bench.py [--gpus N] [--steps K] [--warmup W]          (N > 1: launched under torchrun by the driver)
    python bench.py --impl reference ...                          (the CPU path on the host cores)

Rank 0 prints ONE JSON line.  `value` is whole-job images/s with inputs resident in HBM; `e2e` is the same
path called with HOST (pinned) buffers, host<->device copies inside the timed region; `roofline` is the
dominant kernel against the measured HBM peak; `cpu_baseline` is the CPU path timed on this host.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "refined pseudo-label images/sec (PAR+label+CRF loss, 448^2, 21-cls)"
DILATIONS = [1, 2, 4, 8, 12, 24]
NUM_ITER = 10
THR_HIGH, THR_LOW = 0.7, 0.25
WORKLOADS = {
    "voc": dict(C=21, n_fg=2, H=448, W=448, thr=(0.7, 0.25), name="VOC-shape B=%d/GPU, 3x448x448, 21 classes, 2 fg/img (BASELINE.json configs[1])"),
    "coco": dict(C=81, n_fg=3, H=448, W=448, thr=(0.65, 0.25), name="COCO-shape B=%d/GPU, 3x448x448, 81 classes, 3 fg/img (BASELINE.json configs[2])"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cosa", choices=["cosa", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU (weak scaling)")
    ap.add_argument("--workload", default="voc", choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, default=0, help="override H = W (BASELINE.json configs[3] sweep: 512..1024)")
    ap.add_argument("--aux-labelling", action="store_true",
                    help="also label the auxiliary CAMs of the batch (main.py:171-199, the reference's default); "
                         "reported as a separate workload")
    ap.add_argument("--par-step", default="chain",
                    help="PAR step kernel (cosa_b200.par.set_step_mode): chain[<images per group>] (default), tile, smem")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay measurement")
    ap.add_argument("--no-aux", action="store_true", help="skip the x2-labelling (auxiliary CAMs) side measurement")
    ap.add_argument("--port-only", action="store_true",
                    help="--impl reference: time the oracle port even where /root/reference is mounted")
    args = ap.parse_args()
    if args.aux_labelling:      # a device-side variant only: the host pipeline and the CPU arm run the headline step
        args.no_e2e = True
        args.no_cpu_baseline = True
    return args


# ----------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------
def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the GPU phases of the benchmark run."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        """One nvidia-smi process (rank 0 only) samples every GPU of the job; `gpu_indices` empty = do nothing."""
        self.gpu_indices = list(gpu_indices)
        self.proc = None
        self.file = None

    def __enter__(self):
        if not self.gpu_indices:
            return self
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(str(i) for i in self.gpu_indices),
                                          "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary_peek(self):
        """The clocks seen so far, without ending the sampling."""
        try:
            self.file.flush()
            sm = [float(r.strip().split(", ")[1]) for r in open(self.file.name) if r.count(",") >= 8]
            busy = [v for v in sm if v >= 0.5 * max(sm)] if sm else []
            return {"sm_mhz": statistics.median(busy)} if busy else None
        except Exception:
            return None

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.file is None:
            return out
        try:
            self.file.flush()
            rows = [r.strip().split(", ") for r in open(self.file.name) if r.strip()]
            os.unlink(self.file.name)
            sm = [float(r[1]) for r in rows if len(r) >= 9]
            if sm:
                busy = [v for v in sm if v >= 0.5 * max(sm)] or sm
                out["sm_mhz"] = statistics.median(busy)
                out["sm_max_mhz"] = float(rows[0][2])
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                out["reasons"] = [n for i, n in enumerate(names) if any(r[5 + i].strip() == "Active" for r in rows)]
                out["samples"] = len(sm)
        except Exception as e:   # never let monitoring kill the benchmark
            out["error"] = str(e)
        return out


def make_inputs(args, rank):
    from cosa_b200 import synthetic
    wl = dict(WORKLOADS[args.workload])
    if args.size:
        wl["H"] = wl["W"] = args.size
        wl["name"] = wl["name"].replace("448x448", "%dx%d" % (args.size, args.size)) + " [size sweep]"
    seed = 1000 * (1 if args.workload == "voc" else 2) + rank
    d = synthetic.synthetic_batch(B=args.batch, C=wl["C"], H=wl["H"], W=wl["W"], n_fg=wl["n_fg"], seed=seed)
    d["thr"] = wl["thr"]
    return d, wl


# ----------------------------------------------------------------------------------------------------
# the CPU path (oracle port + the reference C++ lattice when it was built): cpu_baseline and --impl reference
# ----------------------------------------------------------------------------------------------------
def cpu_path_step(d, n_images, full=False):
    """PAR + labelling + CRF loss forward/backward on the first n_images images, on the host cores (the oracle port:
    torch-CPU PAR / labelling + the reference C++ lattice when oracle/_ref was built, else its C port).
    full=True also returns what the parity gate compares: labels, near-tie margins, loss, gradient."""
    from oracle import reference_port as port
    # the reference C++ calls omp_set_num_threads(min(max_threads, N)) (bilateralfilter.cpp:45-47), which also
    # throttles torch's OpenMP pool for everything that follows; give the CPU path all cores again every step
    torch.set_num_threads(os.cpu_count() or 1)
    s = slice(0, n_images)
    cams = port.cam_validation(d["cams"][s], d["cls_label"][s])
    res = port.cam2mask(images=port.denormalize_img(d["simg"][s]), img_boxes=d["img_box"][s], cams=cams,
                        cls_labels=d["cls_label"][s], threshold_high=d["thr"][0], threshold_low=d["thr"][1],
                        refine_model=port.ParOracle(DILATIONS, NUM_ITER), return_margins=full)
    label, margins = res if full else (res, None)
    logit = d["logits"][s].clone().requires_grad_(True)
    loss = port.get_energy_loss(d["simg"][s], logit, label, d["img_box"][s], weight=1e-7, sigma_rgb=15.0,
                                sigma_xy=100.0, scale_factor=0.5)
    loss.backward()
    if full:
        return float(loss.detach()), label, margins, logit.grad, port.last_loss_exact_over_reference()
    return float(loss.detach())


_REF_MODS = None


def reference_path_step(d, n_images):
    """The same step through the REFERENCE's own code (models/PAR.py, utils/seg_helper.py, utils/torch_helper.py and
    the reference C++ lattice), imported unmodified by oracle/reference_import.py.  Only where /root/reference is
    mounted (the build container)."""
    global _REF_MODS
    from oracle import reference_import
    if _REF_MODS is None:
        _REF_MODS = reference_import.load_reference()
    par_mod, sh, th = _REF_MODS
    torch.set_num_threads(os.cpu_count() or 1)
    s = slice(0, n_images)
    with torch.no_grad():
        img_denorm = th.denormalize_img(d["simg"][s])
        cams = sh.cam_validation(d["cams"][s], d["cls_label"][s])
        label = sh.cam2mask(images=img_denorm, img_boxes=d["img_box"][s], cams=cams, cls_labels=d["cls_label"][s],
                            threshold_high=d["thr"][0], threshold_low=d["thr"][1],
                            refine_model=par_mod.PAR(num_iter=NUM_ITER, dilations=DILATIONS))
    logit = d["logits"][s].clone().requires_grad_(True)
    layer = sh.DenseEnergyLoss(weight=1e-7, sigma_rgb=15, sigma_xy=100, scale_factor=0.5)
    loss = sh.get_energy_loss(img=d["simg"][s], logit=logit, label=label, img_box=d["img_box"][s], loss_layer=layer)
    loss.backward()
    return float(loss.detach())


def cpu_kind():
    from oracle import lattice
    return ("port (torch-CPU PAR/labelling port + reference C++ lattice built into oracle/_ref)" if lattice.have_ref()
            else "port (torch-CPU PAR/labelling port + C lattice port)")


REF_SAMPLE_IMAGES = 4      # images per step of the --impl reference arm: the same at every N (one host runs it)
CPU_LEG_IMAGES = 32        # images of the cosa arm's cpu_baseline / parity leg (about 7 s of CPU work at VOC shape)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import reference_import
    torch.set_num_threads(os.cpu_count() or 1)
    d, wl = make_inputs(args, 0)
    use_ref = reference_import.available() and not args.port_only
    step_fn = reference_path_step if use_ref else cpu_path_step
    n_img = min(REF_SAMPLE_IMAGES, args.batch)
    step_fn(d, 1)                                           # import / first-touch warm-up
    for _ in range(args.warmup):
        step_fn(d, n_img)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_fn(d, n_img)
    dt = time.perf_counter() - t0
    value = n_img * args.steps / dt
    cores = os.cpu_count() or 1
    kind = "reference" if use_ref else "port"
    what = ("the reference's own models/PAR.py + utils/seg_helper.py + utils/torch_helper.py + C++ lattice, imported "
            "unmodified (oracle/reference_import.py)" if use_ref else cpu_kind())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"] % args.batch, "par": {"dilations": DILATIONS, "num_iter": NUM_ITER},
                   "crf": "DenseEnergyLoss(1e-7, 15, 100, 0.5)", "sample_images_per_step": n_img,
                   "host": "ONE host runs this arm whatever --gpus says (rank 0; the other ranks exit): the value does "
                           "not scale with N, only the N = 1 ratio compares like with like"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": kind,
                         "sample": "%d images per step of the same seeded batch; %s; torch threads=%d, OpenMP default"
                                   % (n_img, what, cores)},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------
# the CUDA arm
# ----------------------------------------------------------------------------------------------------
PAR_STEP_NOTE = {
    "chain": "1 affinity launch + ONE launch for all 10 propagation steps (north_star's single launch; "
             "par_chain_kernel: one CTA per step x image tile, tile-level step counters instead of grid barriers)",
    "tile": "1 affinity launch + 10 step launches (programmatic dependent launches)",
    "smem": "1 affinity launch + 10 launches of the generic step kernel",
}


def algorithmic_bytes(kernels, B, C, H, W, nc, M):
    """Per-launch algorithmic bytes of every kernel (fp32).  h, w = half resolution."""
    h, w = H // 2, W // 2
    n, K, T, ND = h * w, C, NUM_ITER, 8 * len(DILATIONS)
    Kp = (K + 3) // 4 * 4
    # high + low stacks share the affinity; the last live channel of each stack is derived from the channel sum
    # instead of being propagated (label_kernels.cu: cosa_cam2mask) unless COSA_CAM2MASK_ALL_CHANNELS is set
    ncm = 2 * nc if os.environ.get("COSA_CAM2MASK_ALL_CHANNELS") else 2 * (nc - 1)
    per = {
        # (denormalize_img / cam_validation are folded into cam2mask_prepare on the bench path: these two entries only
        # appear when something else materialises their result)
        "denormalize_img_kernel": 2 * 4 * B * 3 * H * W,
        "cam_validation_kernel": 4 * B * ((C - 1) + (nc - 1)) * H * W,
        "cam2mask_keys_kernel": 4 * B * (C - 1) + 4 * B * C,
        # reads the normalised image and the planes of the present classes, writes the half-res image and the stacks
        "cam2mask_prepare_kernel": 4 * B * (H * W * (3 + (nc - 1)) + n * (3 + ncm)),
        "par_affinity_kernel": 4 * B * n * (3 + ND),
        # one propagation step.  SURVEY 8(d): masks read + written once, affinities on-chip.  This design streams
        # the affinity planes as well (4*B*n*ND more bytes per step); reported separately as "design_bytes".
        "par_iterate_kernel": 4 * B * n * (2 * ncm),
        "cam2mask_finalize_kernel": 4 * B * (n * ncm + H * W),
        "energy_prepare_kernel": 4 * B * (H * W * (K + 2) + n * (3 + K + 2)),
        # SURVEY 8(d) "lattice build": n*(12 + 48) + M*(10 + 48), split over the three kernels that do it: per pixel
        # RGB in, (index, weight) x 6 out | per vertex the key | per vertex the blur-neighbour table
        "lattice_reset_kernel": None,
        "lattice_table_clear_kernel": None,                      # sized by the table, not by the problem
        "lattice_tile_build_kernel": B * n * (12 + 48),
        "lattice_insert_kernel": M * 10,
        "lattice_finish_kernel": M * 48,
        "lattice_zero_values_kernel": None,                      # a no-op on this path (the build leaves the rows zeroed)
        "lattice_splat_kernel": B * n * (48 + 4 * K) + 4 * K * M,
        "lattice_blur_kernel": M * (8 + 8 * K),                  # one axis
        "lattice_slice_kernel": B * n * (48 + 4 * K) + 4 * K * M + 4 * B * n * (K + 1),
        "energy_loss_finalize_kernel": 12,
        "energy_logit_grad_kernel": 4 * B * (2 * H * W * K + n * (K + 1)),
    }
    base = lambda k: (k.replace("_vec_kernel", "_kernel").replace("_reg_kernel", "_kernel").replace("_pair_kernel", "_kernel").replace("_smem_kernel", "_kernel")
                      .replace("_tile_kernel", "_kernel").replace("_x2_kernel", "_kernel"))
    out = {k: per.get(base(k)) for k in kernels}
    design = {k: (4 * B * n * (ND + 2 * ncm) if base(k) == "par_iterate_kernel" else out[k]) for k in kernels}
    for whole in ("par_chain_kernel",):   # every step of a PAR call in one launch
        if whole in kernels:
            steps_per_launch = T / max(1.0, kernels[whole])
            out[whole] = int(per["par_iterate_kernel"] * steps_per_launch)
            design[whole] = int(4 * B * n * (ND + 2 * ncm) * steps_per_launch)
    return out, design


def run_cosa_arm(args):
    import torch.distributed as dist
    import cosa_b200
    from cosa_b200 import _lib, seg_helper, sharding, synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the cosa_b200 path has no CPU fallback "
                         "(use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    host, wl = make_inputs(args, rank)
    B, C, H, W = args.batch, wl["C"], wl["H"], wl["W"]
    pinned = {k: v.pin_memory() for k, v in host.items() if isinstance(v, torch.Tensor) and k != "img_box"}
    boxes = host["img_box"]
    thr_high, thr_low = wl["thr"]
    d = {k: v.to(dev) for k, v in pinned.items()}
    # auxiliary CAMs for the x2-labelling variant: the same blobs seen a little differently (mirrored mix of the batch)
    want_aux = args.aux_labelling or not args.no_aux
    if want_aux:
        d["cams_aux"] = (0.8 * d["cams"] + 0.2 * d["cams"].flip(0).flip(-1)).contiguous()
    if args.aux_labelling:
        wl["name"] += " + auxiliary CAMs labelled too (x2 labelling, main.py:171-199)"
    cosa_b200.par.set_step_mode(args.par_step)
    par = cosa_b200.PAR(num_iter=NUM_ITER, dilations=DILATIONS).to(dev)
    layer = cosa_b200.DenseEnergyLoss(weight=1e-7, sigma_rgb=15, sigma_xy=100, scale_factor=0.5)

    def step(t, overlap=True, aux=args.aux_labelling, img_box=None):
        boxes = host["img_box"] if img_box is None else img_box
        # the CRF lattice needs only the image: DenseEnergyLoss.prebuild_lattice starts it on a second stream - under
        # cam2mask next to the per-step PAR kernel, else behind cam2mask, beside the first kernel of get_energy_loss,
        # which picks it up (same kernels, same results)
        early = cosa_b200.par.lattice_prebuild_before_cam2mask()
        if overlap and early:
            layer.prebuild_lattice(t["simg"], C)
        img_denorm = cosa_b200.denormalize_img(t["simg"])                       # main.py:117
        cams = cosa_b200.cam_validation(t["cams"], t["cls_label"])
        if aux:
            # main.py's default (aux_cam2seg=True, :171-199): the auxiliary CAMs of the batch are labelled as well;
            # the two cam2mask calls share the images, hence the PAR affinity
            with par.shared_affinity():
                label = cosa_b200.cam2mask(images=img_denorm, img_boxes=boxes, cams=cams, cls_labels=t["cls_label"],
                                           threshold_high=thr_high, threshold_low=thr_low, refine_model=par)
                aux = cosa_b200.cam_validation(t["cams_aux"], t["cls_label"])
                cosa_b200.cam2mask(images=img_denorm, img_boxes=boxes, cams=aux, cls_labels=t["cls_label"],
                                   threshold_high=thr_high, threshold_low=thr_low, refine_model=par)
        else:
            label = cosa_b200.cam2mask(images=img_denorm, img_boxes=boxes, cams=cams, cls_labels=t["cls_label"],
                                       threshold_high=thr_high, threshold_low=thr_low, refine_model=par)
        if overlap and not early:
            layer.prebuild_lattice(t["simg"], C)
        logit = t["logits"].detach().requires_grad_(True)
        loss = cosa_b200.get_energy_loss(img=t["simg"], logit=logit, label=label, img_box=boxes, loss_layer=layer)
        loss.backward()
        return label, loss, logit.grad

    def sync_all():
        torch.cuda.synchronize()
        sharding.barrier()
        torch.cuda.synchronize()

    # ---- parity gate + CPU baseline (rank 0, N = 1 only): the CPU path on the first images of this very batch, its
    # labels / loss / gradient compared with the GPU's on the same images BEFORE anything is timed --------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_img = min(CPU_LEG_IMAGES, B)
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_path_step(host, 1)
        t0 = time.perf_counter()
        cpu_loss, cpu_label, cpu_margin, cpu_grad, exact_ratio = cpu_path_step(host, n_img, full=True)
        dt = time.perf_counter() - t0
        cpu = {"value": n_img / dt, "unit": "images/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "first %d images of the same batch, 1 warm-up image + 1 timed pass (%.1f s); %s"
                         % (n_img, dt, cpu_kind())}
        sub = {k: v[:n_img].contiguous() for k, v in d.items()}
        g_label, g_loss, g_grad = step(sub, img_box=host["img_box"][:n_img])
        diff = (g_label.cpu() != cpu_label)
        flips = int(diff.sum())
        worst = float(cpu_margin[diff].max()) if flips else 0.0
        # the GPU's loss / gradient on the CPU's labels (a near-tie flip must not leak into the comparison)
        logit = sub["logits"].detach().requires_grad_(True)
        l2 = cosa_b200.get_energy_loss(img=sub["simg"], logit=logit, label=cpu_label.to(dev),
                                       img_box=host["img_box"][:n_img], loss_layer=layer)
        l2.backward()
        # the reference accumulates <S, AS> with a float32 np.dot (seg_helper.py:890) that is itself ~1e-4 away from
        # the exact sum at these sizes; the product accumulates in double.  Both distances are reported: against the
        # CPU leg's own value, and against the same maths accumulated in float64 (oracle/reference_port.py: LAST_DOT)
        loss_rel = abs(float(l2.detach()) - cpu_loss) / abs(cpu_loss)
        loss_rel_exact = abs(float(l2.detach()) - cpu_loss * exact_ratio) / abs(cpu_loss * exact_ratio)
        drift = abs(exact_ratio - 1.0)
        grad_rel = float((logit.grad.cpu() - cpu_grad).abs().max() / cpu_grad.abs().max())
        parity = {"images": n_img, "label_pixels": int(cpu_label.numel()), "label_flips": flips,
                  "worst_flip_margin": worst, "near_tie_margin": 1e-5, "loss_rel": loss_rel,
                  "loss_rel_vs_float64_accumulation": loss_rel_exact, "cpu_float32_dot_drift": drift,
                  "grad_rel": grad_rel, "tolerance": 1e-4, "against": cpu["kind"]}
        ok = ((flips == 0 or (worst <= 1e-5 and flips <= 1e-5 * cpu_label.numel())) and loss_rel_exact <= 1e-4
              and loss_rel <= drift + 1e-4 and grad_rel <= 1e-4)
        parity["ok"] = bool(ok)
        if not ok:
            emit({"metric": METRIC, "error": "parity gate failed - nothing was timed", "parity": parity})
            raise SystemExit("bench.py: GPU results differ from the CPU path: %s" % json.dumps(parity))
        del sub, logit, l2, g_label, g_loss, g_grad

    # ---- device-resident throughput ------------------------------------------------------------------
    clocks = ClockSampler(range(world) if rank == 0 else [])     # one sampler process for all GPUs of the job
    clocks.__enter__()          # sampled from the warm-up to the end of the e2e loop (nvidia-smi period: 100 ms)
    for _ in range(max(3, args.warmup)):
        label, loss, grad = step(d)
    sync_all()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        label, loss, grad = step(d)
    ev1.record()
    sync_all()
    ms_total = sharding.all_reduce_max(ev0.elapsed_time(ev1))
    launches = _lib.launch_count() - launches0
    total_images = sharding.all_reduce_sum(B * args.steps)
    value = total_images / (ms_total / 1e3)
    mean_loss = sharding.mean_loss_over_ranks(float(loss.detach()), B)

    # ---- the same step replayed from ONE CUDA graph (cosa_b200.GraphedStep): what the GPU needs when the host's
    # launch path is out of the way; matters for small batches (configs[0]) ----------------------------------
    graph_replay = None
    if not args.aux_labelling and not args.no_graph:
        gs = cosa_b200.GraphedStep(par, layer, thr_high, thr_low, B=B, C=C, H=H, W=W, img_box=boxes, device=dev)
        gs.simg.copy_(d["simg"]); gs.cams.copy_(d["cams"]); gs.cls_label.copy_(d["cls_label"])
        with torch.no_grad():
            gs.logits.copy_(d["logits"])
        for _ in range(3):
            gs()
        sync_all()
        ev0.record()
        for _ in range(args.steps):
            gs()
        ev1.record()
        sync_all()
        g_ms = sharding.all_reduce_max(ev0.elapsed_time(ev1))
        assert abs(float(gs.loss.detach()) - float(loss.detach())) <= 1e-5 * abs(float(loss.detach())) + 1e-12
        graph_replay = {"value": total_images / (g_ms / 1e3), "unit": "images/s", "ms_per_step": g_ms / args.steps,
                        "note": "cosa_b200.GraphedStep: the step captured in one CUDA graph, one launch per step"}
        del gs

    # ---- the reference's default step also labels the auxiliary CAMs (main.py:171-203, aux_cam2seg=True): the same
    # step with that second cam2mask call, as a side measurement ("x2 labelling") -------------------
    aux_line = None
    if want_aux and not args.aux_labelling:
        for _ in range(3):
            step(d, aux=True)
        sync_all()
        a_steps = max(3, min(args.steps, 10))
        ev0.record()
        for _ in range(a_steps):
            step(d, aux=True)
        ev1.record()
        sync_all()
        a_ms = sharding.all_reduce_max(ev0.elapsed_time(ev1))
        aux_line = {"value": sharding.all_reduce_sum(B * a_steps) / (a_ms / 1e3), "unit": "images/s",
                    "ms_per_step": a_ms / a_steps, "steps": a_steps,
                    "note": "the step + cam_validation / cam2mask of the auxiliary CAMs (main.py:171-203); the two "
                            "cam2mask calls share the PAR affinity (PAR.shared_affinity)"}

    # ---- per-kernel event timing for the roofline (same inputs, same stream) ---------------------------
    prof_steps = min(args.steps, 5)
    _lib.profile_begin()
    for _ in range(prof_steps):
        step(d, overlap=False)      # one stream: a kernel's event time is its own, not shared with a concurrent one
    prof = _lib.profile_end()
    M_vertices = seg_helper.last_energy_lattice_stats(B, C, H, W, dev)[0]
    nc = 1 + wl["n_fg"]
    peak, peak_src = measured_peak_gbs()
    alg, design = algorithmic_bytes({k: c / prof_steps for k, (c, _) in prof.items()}, B, C, H, W, nc, M_vertices)
    kernels = []
    for name, (count, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        avg_ms = ms / count
        ab = alg.get(name)
        gbs = (ab / 1e9) / (avg_ms / 1e3) if ab else None
        kernels.append({"kernel": name, "launches_per_step": count / prof_steps, "avg_ms": round(avg_ms, 5),
                        "ms_per_step": round(ms / prof_steps, 5), "alg_bytes": ab, "design_bytes": design.get(name),
                        "achieved_gbs": round(gbs, 1) if gbs else None,
                        "frac": round(gbs / peak, 4) if gbs else None})
    top = kernels[0]
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture - only when
    # that capture was made on THIS configuration (profiles/dram_traffic.json names it); null otherwise
    traffic = None
    config_key = "%s_b%d_%dx%d" % (args.workload, B, H, W)
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("config") == config_key:
            traffic = tj.get("kernels", {}).get(top["kernel"])
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": top["kernel"], "achieved": top["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": top["frac"], "traffic": traffic, "peak_source": peak_src,
                "avg_launch_ms": top["avg_ms"], "alg_bytes_per_launch": top["alg_bytes"],
                "share_of_step": round(top["ms_per_step"] / sum(k["ms_per_step"] for k in kernels), 4)}
    if top["kernel"] in ("par_iterate_tile_kernel", "par_chain_kernel"):
        # The PAR step is bound on the SM side, not by HBM: every FMA consumes one neighbour value
        # through the 128 B/clk shared-memory / L1 load path.  Bytes through that path per launch: per pixel and moved
        # channel 45 LDS.128 per quad (180 B), plus the 48 affinity quads (192 B per pixel) once per CTA of a tile
        # (two CTAs share a tile's channels).  Peak = SMs x 128 B/clk x the SM clock seen during the run.
        h2, w2 = H // 2, W // 2
        ncm = 2 * nc if os.environ.get("COSA_CAM2MASK_ALL_CHANNELS") else 2 * (nc - 1)
        if top["kernel"] == "par_chain_kernel":               # <= 4 channels per pass, all steps in the launch
            n_groups, steps_in_launch = (ncm + 3) // 4, NUM_ITER / max(1.0, top["launches_per_step"])
        else:                                                  # channel groups of a tile: each loads the affinity quads
            n_groups, steps_in_launch = (1 if ncm <= 3 else 2 * ((ncm + 5) // 6)), 1
        lsu_bytes = int(B * h2 * w2 * (ncm * 180 + n_groups * 192) * steps_in_launch)
        sm_mhz = (clocks.summary_peek() or {}).get("sm_mhz") or 1965.0
        lsu_peak = torch.cuda.get_device_properties(dev).multi_processor_count * 128 * sm_mhz * 1e6 / 1e9
        lsu_gbs = lsu_bytes / 1e9 / (top["avg_ms"] / 1e3)
        roofline["secondary"] = {"bound": "shared-memory / L1 load path (LSU)", "achieved": round(lsu_gbs, 1),
                                 "peak": round(lsu_peak, 1), "unit": "GB/s", "frac": round(lsu_gbs / lsu_peak, 4),
                                 "bytes_per_launch": lsu_bytes,
                                 "note": "per pixel: 180 B of mask quads per moved channel + 192 B of affinity quads per "
                                         "channel group of its tile, per propagation step; peak = SMs x 128 B/clk x SM clock"}

    # ---- end to end through the host-buffer API (cosa_b200.HostPipeline): pinned host tensors in, labels + loss in
    # pinned host memory out; every step's uploads and read-backs are inside the timed region ---------------------
    e2e = None
    e2e_native = None
    if not args.no_e2e:
        pipe = cosa_b200.HostPipeline(par, layer, threshold_high=thr_high, threshold_low=thr_low, device=dev)
        batch = dict(pinned, img_box=boxes)
        for _ in range(3):
            pipe.submit(batch)
        pipe.drain()
        sync_all()
        e2e_steps = max(3, min(args.steps, 10))
        h2d0, d2h0 = pipe.h2d_bytes, pipe.d2h_bytes
        ev0.record()
        for _ in range(e2e_steps):
            pipe.submit(batch)
        res = pipe.drain()
        ev1.record()
        sync_all()
        e2e_ms = sharding.all_reduce_max(ev0.elapsed_time(ev1))
        ref_loss = float(loss.detach())
        assert abs(float(res[-1][1].detach()) - ref_loss) <= 1e-6 * abs(ref_loss) + 1e-12, "e2e loss differs from the device path"
        e2e = {"value": sharding.all_reduce_sum(B * e2e_steps) / (e2e_ms / 1e3), "unit": "images/s",
               "h2d_bytes_per_step": (pipe.h2d_bytes - h2d0) // e2e_steps,
               "d2h_bytes_per_step": (pipe.d2h_bytes - d2h0) // e2e_steps, "steps": e2e_steps,
               "ms_per_step": e2e_ms / e2e_steps,
               "note": "cosa_b200.HostPipeline: pinned host buffers; upload of step i+1 overlaps the kernels of step i "
                       "(copy stream); CAM planes of absent classes are not uploaded (zero after cam_validation), the [0,1] image "
                       "is derived on the device from the normalised one as in main.py:117; "
                       "labels + loss read back every step"}

        # the same step fed one stage further upstream, where the tensors are small: the teacher's raw multi-scale
        # CAMs and the decoder's logits on the ViT token grids (merge + validation and the enlargement run on the device)
        raw_cams = [t.pin_memory() for t in synthetic.synthetic_raw_cams(host, seed=7000 + rank)]
        nbatch = dict(simg=pinned["simg"], raw_cams=raw_cams, seg_lowres=pinned["seg_lowres"],
                      cls_label=pinned["cls_label"], img_box=boxes)
        npipe = cosa_b200.HostPipeline(par, layer, threshold_high=thr_high, threshold_low=thr_low, device=dev)
        for _ in range(3):
            npipe.submit_native(nbatch)
        npipe.drain()
        sync_all()
        h2d0, d2h0, l0 = npipe.h2d_bytes, npipe.d2h_bytes, _lib.launch_count() + npipe.graph_kernels
        ev0.record()
        for _ in range(e2e_steps):
            npipe.submit_native(nbatch)
        npipe.drain()
        ev1.record()
        sync_all()
        n_ms = sharding.all_reduce_max(ev0.elapsed_time(ev1))
        e2e_native = {"value": sharding.all_reduce_sum(B * e2e_steps) / (n_ms / 1e3), "unit": "images/s",
                      "h2d_bytes_per_step": (npipe.h2d_bytes - h2d0) // e2e_steps,
                      "d2h_bytes_per_step": (npipe.d2h_bytes - d2h0) // e2e_steps, "steps": e2e_steps,
                      "ms_per_step": n_ms / e2e_steps,
                      "gpu_launches_per_step": (_lib.launch_count() + npipe.graph_kernels - l0) / e2e_steps,
                      "note": "HostPipeline.submit_native: host buffers at the resolution the networks emit them "
                              "(raw CAMs on the 28/14/42 token grids for images and flips, logits 28x28); "
                              "multi_scale_cam_merge + cam_validation and the main.py:167 enlargement (with its "
                              "adjoint in the backward) run on the device in front of the same step"}

    clocks.__exit__(None, None, None)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"] % B, "par": {"dilations": DILATIONS, "num_iter": NUM_ITER},
                       "crf": "DenseEnergyLoss(1e-7, 15, 100, 0.5)", "thresholds": [thr_high, thr_low],
                       "par_step": PAR_STEP_NOTE[args.par_step.rstrip("0123456789")],
                       "fused_producers": "denormalize_img and cam_validation are folded into cam2mask's first kernel "
                                          "(cosa_cam2mask_ex); their tensors are never written",
                       "streams": ("single stream (COSA_NO_PREBUILD=1)" if os.environ.get("COSA_NO_PREBUILD") == "1" else
                                   "CRF lattice build (image-only) on a second stream (DenseEnergyLoss.prebuild_lattice): "
                                   + ("under cam2mask" if cosa_b200.par.lattice_prebuild_before_cam2mask() else
                                      "beside the softmax / gate kernel of get_energy_loss")),
                       "parallelism": "batch-sharded, %d image(s)/GPU x %d GPU, no data-path collective" % (B, world),
                       "cache": "inputs per step (%.0f MB) exceed the 126 MB L2; no explicit flush"
                                % ((sum(v.numel() * v.element_size() for v in d.values())) / 1e6),
                       "lattice_vertices": M_vertices, "lattice_M_over_n": round(M_vertices / (B * (H // 2) * (W // 2)), 4)},
            "clocks": clocks.summary(), "e2e": e2e, "e2e_native": e2e_native, "graph_replay": graph_replay,
            "aux_labelling": aux_line, "gpu_launches": launches, "roofline": roofline, "parity": parity,
            "cpu_baseline": cpu, "kernels": kernels, "loss": mean_loss,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line.  Native libraries write to file descriptor 1 behind Python's back (NCCL
    prints its version banner there when NCCL_DEBUG=VERSION is in the environment, and ignores NCCL_DEBUG_FILE at
    that level), so the descriptor is pointed at stderr for the rest of the process and the JSON line goes to a
    private duplicate of the original stdout."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    print(json.dumps(line), file=out, flush=True)


def main():
    args = parse_args()
    _claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_cosa_arm(args)


if __name__ == "__main__":
    main()
