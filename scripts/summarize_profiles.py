"""Turn what a gpurun profiling call brought back (gpurun_out/) into the small tracked summaries under profiles/.

  python scripts/summarize_profiles.py launches gpurun_out/launches_r1.csv profiles/r1_launches_r50_b8.md [step_index]
  python scripts/summarize_profiles.py ncu gpurun_out/conv_tower.ncu-rep profiles/r1_ncu_conv_tower.md
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__inst_executed.sum.per_cycle_elapsed",
]


def launches(src, dst, step=3):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in rows]
    starts = [i for i, (n, _) in enumerate(names) if "preprocess_kernel" in n]
    # one step = [set_ints, memset, preprocess ... finalize]; the two launches before preprocess belong to it
    a, b = starts[step] - 2, starts[step + 1] - 2
    agg = collections.OrderedDict()
    for n, v in names[a:b]:
        k = re.sub(r"\(.*", "", n).replace("void ", "").replace("dafne::", "")
        agg.setdefault(k, [0.0, 0])
        agg[k][0] += v
        agg[k][1] += 1
    tot = sum(v[0] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list, one step (launches {a}..{b - 1} of {len(names)}), source `{src}`\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES)\n\n")
        f.write(f"total {tot / 1e6:.3f} ms over {b - a} launches\n\n| kernel | launches | us | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"| `{k}` | {v[1]} | {v[0] / 1e3:.1f} | {100 * v[0] / tot:.1f}% |\n")
    print(open(dst).read())


def ncu(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of `{src}`\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"## {d.get('Kernel Name', '?')[:120]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
            f.write("\n")
    print(open(dst).read())


def tensor_pipe(csv_path, per_launch_json, dst, first_global_index):
    """ncu per-launch tensor-pipe utilisation of the conv / stem kernels of one forward, labelled with the layer names
    of bench.py's per-launch profile (same launch order), FLOP-weighted per group."""
    import json

    rows = [l for l in open(csv_path) if not l.startswith("==")]
    by = collections.OrderedDict()
    for x in csv.DictReader(rows):
        by.setdefault(int(x["ID"]), {"kernel": x["Kernel Name"]})[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
    ops = [o for o in json.load(open(per_launch_json)) if o["kind"] in (1, 2)]
    launches_per_forward = len(ops)
    seen = {}
    for k, m in by.items():
        op = (first_global_index + k) % launches_per_forward
        seen.setdefault(op, m)
    groups = collections.OrderedDict()
    lines = []
    for i, o in enumerate(ops):
        m = seen.get(i)
        if m is None:
            continue
        nm = o["name"]
        g = ("head towers" if "tower" in nm else "head predictions" if nm.startswith("pred") else
             "FPN + P6/P7" if ("fpn" in nm or "top_block" in nm) else "stem" if nm.startswith("stem") else "ResNet " + nm.split("bottom_up.")[-1].split(".")[0])
        tp = m["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
        us = m["gpu__time_duration.sum"] / 1e3
        dram = (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / 1e6
        a = groups.setdefault(g, [0.0, 0.0, 0.0, 0.0, 0])
        a[0] += o["flops"]
        a[1] += o["flops"] * tp
        a[2] += us
        a[3] += dram
        a[4] += 1
        lines.append(f"| `{nm[-40:]}` | {m['kernel'].split('(')[0].replace('void ', '')} | {us:.1f} | {tp:.1f} | "
                     f"{o['flops'] / us / 1e6:.0f} | {dram:.1f} | {m['dram__throughput.avg.pct_of_peak_sustained_elapsed']:.0f} |")
    with open(dst, "w") as f:
        f.write(f"# Tensor-pipe utilisation per convolution launch, one forward (source `{csv_path}`)\n\n"
                "`ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__bytes_*` "
                "(`--clock-control none`; serialised, cold caches: durations are upper bounds)\n\n"
                "## FLOP-weighted per group\n\n| group | launches | GFLOP | us (ncu) | tensor pipe % (FLOP-weighted) | TFLOP/s | DRAM MB |\n|---|---|---|---|---|---|---|\n")
        tot = [0.0, 0.0, 0.0]
        for g, a in groups.items():
            f.write(f"| {g} | {a[4]} | {a[0] / 1e9:.1f} | {a[2]:.0f} | {a[1] / a[0]:.1f} | {a[0] / a[2] / 1e6:.0f} | {a[3]:.0f} |\n")
        bb = [a for g, a in groups.items() if g.startswith("ResNet") or g.startswith("FPN")]
        if bb:
            fl = sum(a[0] for a in bb)
            f.write(f"| **backbone + FPN** | {sum(a[4] for a in bb)} | {fl / 1e9:.1f} | {sum(a[2] for a in bb):.0f} | "
                    f"**{sum(a[1] for a in bb) / fl:.1f}** | {fl / sum(a[2] for a in bb) / 1e6:.0f} | {sum(a[3] for a in bb):.0f} |\n")
        f.write("\n## Per launch\n\n| layer | kernel | us | tensor pipe % | TFLOP/s | DRAM MB | DRAM % of peak |\n|---|---|---|---|---|---|---|\n")
        f.write("\n".join(lines) + "\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 3)
    elif sys.argv[1] == "tensor_pipe":
        tensor_pipe(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]))
    else:
        ncu(sys.argv[2], sys.argv[3])
