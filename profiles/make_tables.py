#!/usr/bin/env python
"""Markdown tables of profiles/README.md from the bench lines of one evidence set.

    python profiles/make_tables.py r02n [r02m]      # tag of the 1-GPU set, tag of the multi-GPU lines

Prints: headline table, scaling table, per-kernel table, algorithmic bytes against DRAM traffic.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    with open(os.path.join(HERE, name)) as f:
        return json.load(f)


def main(tag, mtag):
    d = load("%s_bench_voc_b32.json" % tag)
    co, b4 = load("%s_bench_coco.json" % tag), load("%s_bench_b4.json" % tag)
    sz = [load("%s_bench_sz%d.json" % (tag, s)) for s in (512, 768, 1024)]
    ti, ax = load("%s_bench_voc_b32_tile.json" % tag), load("%s_bench_voc_b32_aux_labelling.json" % tag)
    rf, fs = load("%s_bench_reference.json" % tag), load("%s_bench_full_step.json" % tag)
    print("| Config | images/s (device) | ms/step | end to end, full-resolution pinned host buffers | end to end, "
          "network-resolution host buffers | CPU path on the 16 host cores |\n|---|---|---|---|---|---|")
    print("| VOC B = 32, 448², 21 cls (`%s_bench_voc_b32.json`) | **%.0f** | **%.3f** (%.3f replayed from one CUDA graph) "
          "| %.0f (%.1f ms/step, %.0f MB H2D per step: PCIe-bound) | %.0f (%.2f ms/step, %.0f MB H2D) | %.2f on %d images "
          "(`--impl reference`: %.2f) |" % (tag, d["value"], d["ms_per_step"], d["graph_replay"]["ms_per_step"],
                                            d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"] / 1e6,
                                            d["e2e_native"]["value"], d["e2e_native"]["ms_per_step"],
                                            d["e2e_native"]["h2d_bytes_per_step"] / 1e6, d["cpu_baseline"]["value"],
                                            d["parity"]["images"], rf["value"]))
    print("| + the auxiliary CAMs labelled too | %.0f | %.2f | | | |" % (ax["value"], ax["ms_per_step"]))
    print("| the same with the per-step PAR kernel (\"tile\") | %.0f | %.3f | | | |" % (ti["value"], ti["ms_per_step"]))
    print("| COCO shape, 81 cls, 3 fg, B = 32 | %.0f | %.2f | %.0f | | %.2f |"
          % (co["value"], co["ms_per_step"], co["e2e"]["value"], co["cpu_baseline"]["value"]))
    print("| VOC B = 4 (configs[0]) | %.0f | %.3f (%.3f from the graph) | %.0f | | |"
          % (b4["value"], b4["ms_per_step"], b4["graph_replay"]["ms_per_step"], b4["e2e"]["value"]))
    print("| 512² / 768² / 1024², B = 4 | %s | %s | | | |" % (" / ".join("%.0f" % s["value"] for s in sz),
                                                             " / ".join("%.2f" % s["ms_per_step"] for s in sz)))
    print("\nconfigs[4]: %.0f images/s, %.1f ms per step, path share %.1f %%\n"
          % (fs["value"], fs["ms_per_step"], 100 * fs["path_share_of_step"]))

    base = load("%s_bench_voc_b32.json" % mtag) if os.path.exists(os.path.join(HERE, "%s_bench_voc_b32.json" % mtag)) else d
    print("| GPUs | images/s device-resident | ms/step | replayed from the CUDA graph | end to end, full-resolution host "
          "buffers | end to end, network-resolution host buffers |\n|---|---|---|---|---|---|")
    for n in (1, 2, 4, 8):
        name = "%s_bench_voc_b32.json" % tag if n == 1 else "%s_bench_voc_b32_%dgpu.json" % (mtag, n)
        if not os.path.exists(os.path.join(HERE, name)):
            continue
        m = load(name)
        ref = d if n == 1 else base
        tot = lambda e: e["value"] / 32 * (e["h2d_bytes_per_step"] + e["d2h_bytes_per_step"]) / 1e9
        print("| %d | %.0f (%.2f×) | %.3f | %.0f (%.2f×) | %.0f (%.2f×, %.0f GB/s over PCIe in total) | %.0f (%.2f×, %.0f GB/s) |"
              % (n, m["value"], m["value"] / ref["value"], m["ms_per_step"], m["graph_replay"]["value"],
                 m["graph_replay"]["value"] / ref["graph_replay"]["value"], m["e2e"]["value"],
                 m["e2e"]["value"] / ref["e2e"]["value"], tot(m["e2e"]), m["e2e_native"]["value"],
                 m["e2e_native"]["value"] / ref["e2e_native"]["value"], tot(m["e2e_native"])))

    total = sum(k["ms_per_step"] for k in d["kernels"])
    print("\n| Kernel | launches/step | ms/step | share | frac of HBM peak |\n|---|---|---|---|---|")
    for k in d["kernels"]:
        print("| `%s` | %g | %.3f | %.0f %% | %s |" % (k["kernel"], k["launches_per_step"], k["ms_per_step"],
                                                       100 * k["ms_per_step"] / total, ("%.2f" % k["frac"]) if k["frac"] else ""))
    tr = load("dram_traffic.json")["kernels"]
    print("\n| Kernel | algorithmic MB | design MB | DRAM traffic MB | traffic ÷ algorithmic |\n|---|---|---|---|---|")
    for k in d["kernels"]:
        key = k["kernel"].replace("cam2mask_prepare_kernel", "cam2mask_prepare_x2_kernel")
        if key in tr and k["alg_bytes"]:
            print("| `%s` | %.0f | %.0f | %.0f | %.2f |" % (k["kernel"], k["alg_bytes"] / 1e6,
                                                            (k["design_bytes"] or k["alg_bytes"]) / 1e6, tr[key] / 1e6,
                                                            tr[key] / k["alg_bytes"]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
