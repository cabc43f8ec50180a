#!/usr/bin/env python
"""Benchmark of the DAFNe inference hot path on B200 (contract: see the task description / DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload r101_b32|r50_b8|...]

Default workload: `r101_b32` = BASELINE.json configs[2] (ResNet-101 dota-1.0_r101_ms, 32 x 1024^2 on one B200), the
configuration the north_star's targets are quoted on; under torchrun (WORLD_SIZE > 1) `r101_b16` = configs[3]
(16 images per GPU, global batch 128 on 8 GPUs). The other names are alternates for profiles/.

A "step" = one pass of the hot path (normalise -> ResNet+FPN -> head -> threshold/top-k/decode/rotated NMS) over one
batch of synthetic 1024x1024x3 uint8 images. `value` = images/s with the inputs already resident in HBM; `e2e` = the
same through the reference-facing C-ABI call with HOST buffers, `dafne_detect_host` in its pipelined form
(`dafne_detect_host_begin` / `_end`, two batches in flight): H2D of every step's images and D2H of its detections are
inside the timed region, the copy of step i+1 overlapping the compute of step i. Under torchrun every rank runs its own shard of the batch (weak scaling) and a
step ends with ONE all-gather of the fixed-shape detection record (dafne_b200.distributed.gather_wire), straight from
device memory in both legs.

`--impl reference` times the reference's CPU implementation of the same path on the host cores. detectron2 / poly_nms
cannot be installed here (no network), so that arm is the restated oracle (kind "port"): torch CPU fp32 ops +
the C polygon NMS, all host threads, one image per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (config file, per-GPU batch, H, W, BASELINE.json config it restates)
    "r50_b8": ("configs/dota10_r50_1024.yaml", 8, 1024, 1024, "configs[1]: ResNet-50 + FPN DAFNe, batch 8x1024x1024"),
    "r101_b32": ("configs/dota10_r101_ms.yaml", 32, 1024, 1024, "configs[2]: ResNet-101 DAFNe dota-1.0_r101_ms, batch 32"),
    "r101_b16": ("configs/dota10_r101_ms.yaml", 16, 1024, 1024, "configs[3]: ResNet-101 DAFNe, 16 images per GPU"),
    # SURVEY 8(d) worst case for the rotated NMS: the class bias is raised until the per-level top-k cap (2000) binds at
    # p3, p4 and p5, so about 8-9 thousand boxes per image enter the NMS
    "r101_b32_nmsmax": ("configs/dota10_r101_ms.yaml", 32, 1024, 1024,
                        "configs[2] with the NMS worst case of SURVEY 8(d): top-k cap binding at p3/p4/p5"),
    "r50_b8_nmsmax": ("configs/dota10_r50_1024.yaml", 8, 1024, 1024,
                      "configs[1] with the NMS worst case of SURVEY 8(d): top-k cap binding at p3/p4/p5"),
    "hrsc_r50_b8": ("configs/hrsc_r50_ms.yaml", 8, 1024, 1024, "configs[4] (single size): HRSC r50_ms, batch 8"),
    # configs[4] as the reference runs it: images of 512^2 / 800^2 / 1024^2 in ONE batch, zero-padded to the largest
    # (ImageList.from_tensors) -- the padded area is computed like the reference computes it, images/s counts images
    "hrsc_r50_mixed": ("configs/hrsc_r50_ms.yaml", 24, 1024, 1024,
                       "configs[4]: HRSC r50_ms, mixed 512/800/1024 batch of 24 padded to 1024x1024"),
    # the same 24 images grouped by size into three batches of 8 (one engine plan per size); NOT the reference's batching
    # (GroupNorm never sees zero padding here), so no parity claim -- throughput only
    "hrsc_r50_bucketed": ("configs/hrsc_r50_ms.yaml", 24, 1024, 1024,
                          "configs[4] bucketed by size: HRSC r50_ms, 8x512^2 + 8x800^2 + 8x1024^2 per step"),
}
MIXED_SIZES = [(512, 512), (800, 800), (1024, 1024)]
# HRSC (one class) stress recipe: denser candidates (about 14 % of the p3 locations above the threshold instead
# of 1 % of the (location, class) pairs) and 8:1 elongated base quads, so the rotated NMS sees thousands of overlapping
# slender boxes per image even with C = 1.
HRSC_SYNTH = dict(cls_bias=-2.9, base_quad=(-4.0, -0.5, 4.0, -0.5, 4.0, 0.5, -4.0, 0.5))


NMSMAX_CLS_BIAS = {"r101_b32_nmsmax": -4.6, "r50_b8_nmsmax": -2.6}  # calibrated: candidates per level in the bench line


def synth_weights(workload, spec):
    from dafne_b200.weights import synthetic_state_dict

    if workload in NMSMAX_CLS_BIAS:
        return synthetic_state_dict(spec, 0, cls_bias=NMSMAX_CLS_BIAS[workload])
    return synthetic_state_dict(spec, 0, **(HRSC_SYNTH if workload.startswith("hrsc") else {}))
GFLOP_PER_IMAGE = {"r50": 505.97, "r101": 661.13}  # BASELINE.md section 2 (C = 15, 1024^2)


def load_spec(workload):
    from dafne_b200.config import get_cfg
    from dafne_b200.spec import ModelSpec

    cfg_file, batch, H, W, what = WORKLOADS[workload]
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(ROOT, cfg_file))
    return cfg, ModelSpec.from_cfg(cfg), batch, H, W, what


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe). NVML in-process (one
    sample every 5 ms: the timed region of a default run is a quarter of a second) with nvidia-smi as the fallback."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReason* bits
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index: int):
        self.index = index
        self.rows = []  # (sm_mhz, max_mhz, set(reasons))
        self.source = None
        self._stop = threading.Event()
        self._t = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _run_nvml(self):
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        self.source = "nvml"
        while not self._stop.is_set():
            sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            mask = int(get_reasons(h))
            self.rows.append((sm, mx, {k for k, b in self.BITS.items() if mask & b}))
            self._stop.wait(0.005)
        nv.nvmlShutdown()

    def _run_smi(self):
        self.source = "nvidia-smi"
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self._physical_index()}", f"--query-gpu={self.Q}",
                     "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    r = [c.strip() for c in out.split(",")]
                    names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
                    self.rows.append((float(r[0]), float(r[1]),
                                      {n for n, v in zip(names, r[3:7]) if v.lower().startswith("active")}))
            except Exception:
                pass
            self._stop.wait(0.2)

    def _run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = sorted(r[0] for r in self.rows)
        mx = max((r[1] for r in self.rows), default=0.0)
        reasons = set().union(*[r[2] for r in self.rows]) if self.rows else set()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def measured_peaks():
    """(sustained bf16 TFLOP/s, burst bf16 TFLOP/s, HBM GB/s, where from). The tower kernels are timed per launch inside
    a step that runs for tens of milliseconds under the power cap: `frac` is against the sustained peak, `frac_burst`
    against the burst one -- both are reported."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return (p.get("bf16_tflops_sustained", 1383.4), p.get("bf16_tflops", 1651.8), p.get("hbm_gbs", 6555.5),
                "measured (MEASURED_PEAKS.json)")
    return 1400.0, 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_step(sd, spec, images_u8, num_threads):
    """One image through the restated reference on the host cores (oracle = the checker, timed as the CPU baseline)."""
    import torch

    from oracle import model as omodel
    from oracle import postprocess as opost

    torch.set_num_threads(num_threads)
    batch, sizes = omodel.preprocess(list(images_u8), spec.pixel_mean, spec.pixel_std)
    out = omodel.forward_dense(sd, spec.resnet_depth, batch, "fp32")
    res = opost.postprocess([t.numpy() for t in out["logits"]], [t.numpy() for t in out["reg"]],
                            [t.numpy() for t in out["ctr"]], spec.fpn_strides, sizes, None,
                            score_thresh=spec.score_thresh, pre_nms_topk=spec.pre_nms_topk,
                            nms_thresh=spec.nms_thresh, post_nms_topk=spec.post_nms_topk,
                            sort_corners=spec.sort_corners, thresh_with_ctr=spec.thresh_with_ctr,
                            vehicle_merge=spec.vehicle_merge)
    return sum(len(r["scores"]) for r in res)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch

    cfg, spec, batch, H, W, what = load_spec(args.workload)
    cores = os.cpu_count() or 1
    sd = synth_weights(args.workload, spec)
    g = torch.Generator().manual_seed(1234)
    imgs = torch.randint(0, 256, (1, 3, H, W), dtype=torch.uint8, generator=g)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_step(sd, spec, imgs, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(sd, spec, imgs, cores)
    dt = time.perf_counter() - t0
    value = args.steps / dt
    line = {
        "impl": "reference", "metric": "images_per_sec_1024x1024", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same workload as the GPU arm's line; a step of this arm is a bounded SAMPLE of it (one image of the batch),
        # images/s is per image either way
        "config": {"workload": args.workload, "what": what, "per_gpu_batch": batch, "global_batch": batch,
                   "image": [3, H, W], "resnet_depth": spec.resnet_depth, "num_classes": spec.num_classes,
                   "weights": "seeded random init (dafne_b200.weights.synthetic_state_dict)",
                   "parallelism": f"{cores} host threads", "per_step": "1 image (bounded sample of the batch)"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} x 1 image 1024x1024, restated reference (torch CPU fp32 + C polygon NMS)"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------- GPU arm
def claim_stdout():
    """stdout carries exactly ONE JSON line: keep a private handle to the real stdout for it and point fd 1 at stderr for
    everything else in the process (NCCL's banner when NCCL_DEBUG is set on the box, library warnings)."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def run_ours(args):
    out_stream = claim_stdout()
    import torch
    import torch.distributed as dist

    from dafne_b200.distributed import gather_wire
    from dafne_b200.engine import DET, DafneEngine, DetectionWire

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg, spec, batch, H, W, what = load_spec(args.workload)
    if args.batch:
        batch = args.batch
    sd_all = synth_weights(args.workload, spec)
    cap = spec.post_nms_topk + 24
    # a step is a list of buckets: one (batch, H, W) for every workload except the size-bucketed one
    if args.workload.endswith("_bucketed"):
        shapes = [(batch // len(MIXED_SIZES), h, w) for (h, w) in MIXED_SIZES]
    else:
        shapes = [(batch, H, W)]
    g = torch.Generator().manual_seed(1234 + rank)
    step_bytes = sum(b * 3 * h * w for b, h, w in shapes)
    # inputs: several distinct batches so the input stream (not only the ~GB of activations) exceeds the 126 MB L2
    n_sets = max(2, (160 * 2**20) // step_bytes + 1)
    buckets = []
    for b, h, w in shapes:
        e = DafneEngine(spec, dev)
        e.load_state_dict(sd_all)
        bk = {"eng": e, "batch": b, "sizes": [(h, w)] * b}
        if args.workload.endswith("_mixed"):
            bk["sizes"] = [MIXED_SIZES[i % len(MIXED_SIZES)] for i in range(b)]
        bk["host_sets"] = [torch.randint(0, 256, (b, 3, h, w), dtype=torch.uint8, generator=g).pin_memory()
                           for _ in range(n_sets)]
        bk["dev_sets"] = [t.to(dev) for t in bk["host_sets"]]
        # two result buffers: the reference-facing host call is used in its pipelined form (two batches in flight)
        bk["host_dets"] = [torch.empty(b, cap, DET, dtype=torch.float32).pin_memory() for _ in range(2)]
        bk["host_counts"] = [torch.empty(b, dtype=torch.int32).pin_memory() for _ in range(2)]
        # the step's result record, allocated once: detections + counts in ONE buffer (no per-step allocation, and
        # what a rank contributes to the all-gather)
        bk["wire"] = DetectionWire(b, cap, dev)
        if world > 1:
            n_wire = bk["wire"].buf.numel()
            bk["gathered"] = [torch.empty(world, n_wire, dtype=torch.int32, device=dev) for _ in range(2)]
            bk["gathered_host"] = [torch.empty(world, n_wire, dtype=torch.int32).pin_memory() for _ in range(2)]
        buckets.append(bk)
    eng = buckets[-1]["eng"]  # the largest shape: per-launch profile and roofline
    sizes, host_sets, dev_sets = buckets[-1]["sizes"], buckets[-1]["host_sets"], buckets[-1]["dev_sets"]

    side = torch.cuda.Stream(device=dev) if world > 1 else None  # D2H of the gathered record (e2e leg)

    def step_device(i):
        out = None
        for bk in buckets:
            dets, counts = bk["eng"].detect(bk["dev_sets"][i % n_sets], bk["sizes"], None, True, cap, out=bk["wire"])
            if world > 1:  # the path's one exchange: ONE all-gather of every rank's fixed-shape record
                gather_wire(bk["wire"].buf, bk["batch"], cap, DET, out=bk["gathered"][0])
            out = (dets, counts)
        return out

    def host_loop(steps):
        """K steps through dafne_detect_host_begin / _end with HOST buffers: the H2D copy of the next batch (copy
        stream) overlaps the compute of the current one; every batch's images go H2D and its detections come back
        D2H. At most two batches are in flight (and never two of one engine's result buffer)."""
        prev = None
        for i in range(steps):
            for bk in buckets:
                k = i % 2
                t = bk["eng"].detect_host_begin(bk["host_sets"][i % n_sets], bk["sizes"], None, bk["host_dets"][k],
                                                bk["host_counts"][k], cap)
                if world > 1:
                    # the exchange reads the batch's record where the kernels left it (device memory, compute stream,
                    # right behind the last kernel); every rank then reads the gathered record back on a side stream
                    torch.cuda.current_stream().wait_stream(side)  # gathered[k] of two steps ago has been read out
                    gather_wire(bk["eng"].slot_wire(t), bk["batch"], cap, DET, out=bk["gathered"][k])
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        bk["gathered_host"][k].copy_(bk["gathered"][k], non_blocking=True)
                if prev is not None:
                    finish_host(*prev)
                prev = (bk, t, k)
        finish_host(*prev)

    def finish_host(bk, ticket, k):
        bk["eng"].detect_host_end(ticket)  # detections of that batch are now in host_dets[k] / host_counts[k]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()  # all streams of the device, the side stream included

    def timed(fn, steps, whole_loop=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if whole_loop:
            fn(steps)
        else:
            for i in range(steps):
                fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for i in range(max(args.warmup, 3)):
        step_device(i)
    barrier()
    for bk in buckets:
        bk["eng"].stats(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(step_device, args.steps)
    launches, flops = 0, 0.0
    for bk in buckets:
        l_, f_ = bk["eng"].stats(reset=True)
        launches, flops = launches + l_, flops + f_
    host_loop(2)
    ms_host = timed(host_loop, args.steps, whole_loop=True)
    clocks = sampler.stop() if rank == 0 else None
    dets, counts = step_device(0)
    torch.cuda.synchronize()
    det_counts = counts.cpu().tolist()
    post_counts = eng.post_counts()
    nms_work = eng.nms_stats()

    # per-launch timing of one forward (CUDA events on the launching stream) for the roofline of the dominant kernel
    prof = eng.profile_forward(dev_sets[0], sizes)
    dom = [o for o in prof if o["kind"] == 1 and o["block_n"] == 256 and o["ksize"] == 3 and o["cin"] == 256
           and o["cout"] == 256 and o["stride"] == 1 and "tower" in o["name"]]
    peak_tf, peak_burst, peak_hbm, peak_src = measured_peaks()
    dom_ms = sum(o["ms"] for o in dom)
    dom_fl = sum(o["flops"] for o in dom)
    total_ms = sum(o["ms"] for o in prof)
    conv_ms = sum(o["ms"] for o in prof if o["kind"] == 1)
    conv_fl = sum(o["flops"] for o in prof if o["kind"] == 1)
    achieved = dom_fl / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            per = json.load(f).get("per_workload", {})
        # the capture of the workload with this (depth, per-GPU batch, image size); null if none was taken
        key = {(101, 32): "r101_b32", (50, 8): "r50_b8"}.get((spec.resnet_depth, batch))
        if key in per and (H, W) == (1024, 1024):
            traffic = per[key].get("dram_bytes_per_launch")

    line = None
    if rank == 0:
        images = world * batch * args.steps
        value = images / (ms * 1e-3)
        e2e = images / (ms_host * 1e-3)
        depth_key = "r101" if spec.resnet_depth == 101 else "r50"
        gflop_img = flops / (args.steps * batch) / 1e9 if flops else GFLOP_PER_IMAGE[depth_key]
        line = {
            "metric": "images_per_sec_1024x1024", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {
                "workload": args.workload, "what": what, "per_gpu_batch": batch, "global_batch": world * batch,
                "image": [3, H, W], "batches_per_step": [[b, 3, h, w] for b, h, w in shapes], "resnet_depth": spec.resnet_depth, "num_classes": spec.num_classes,
                "weights": "seeded random init (dafne_b200.weights.synthetic_state_dict"
                           + (", HRSC stress recipe)" if args.workload.startswith("hrsc") else ")"),
                "parallelism": f"batch-sharded x{world}, ONE all-gather of the detection record per step (both legs, "
                               f"from device memory)" if world > 1 else "single GPU",
                "l2": f"{n_sets} distinct input sets ({n_sets * step_bytes / 2**20:.0f} MiB) rotate; "
                      f"activation workspace {eng.workspace_bytes / 2**20:.0f} MiB >> 126 MiB L2",
                "detections_per_image": det_counts[:8],
                "nms_boxes_in_per_image": [c["nms_in"] for c in post_counts[:8]],
                "nms_boxes_kept_per_image": [c["nms_kept"] for c in post_counts[:8]],
                "candidates_per_level_image0": post_counts[0]["candidates"],
                "nms_work_per_batch": nms_work,
                "gflop_per_image": gflop_img,
                "conv_roofline_frac_whole_step": value / world * gflop_img * 1e9 / (peak_tf * 1e12),
            },
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": step_bytes,
                    "d2h_bytes_per_step": (batch * cap * DET * 4 + batch * 4) * (1 + (world if world > 1 else 0)),
                    "ms_per_step": ms_host / args.steps,
                    "what": "dafne_detect_host_begin/_end with pinned host buffers, per rank: H2D of the step's images, "
                            "D2H of its detections" + ("; the gathered record of all ranks is read back too"
                                                       if world > 1 else "")},
            "gpu_launches": launches,
            "roofline": {
                "kernel": "conv_tc_kernel<256> (head tower 3x3 256->256 convs, all levels)", "bound": "tensor",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_burst": peak_burst, "frac_burst": achieved / peak_burst, "traffic": traffic,
                "peak_source": peak_src + ": `peak` = sustained bf16 (the kernel is timed inside a long power-capped "
                               "step), `peak_burst` = burst bf16", "launches": len(dom), "ms_in_step": dom_ms,
                "share_of_forward": dom_ms / total_ms if total_ms else None,
                "all_convs": {"achieved": conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms else 0.0,
                              "ms": conv_ms, "forward_ms_sum_of_launches": total_ms},
            },
        }
        if args.profile_json:
            with open(args.profile_json, "w") as f:
                json.dump(prof, f, indent=1)
    # CPU baseline (rank 0, N = 1 only): bounded sample of the same workload on the host cores
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sd = sd_all
        n_img = args.cpu_images
        t0 = time.perf_counter()
        for k in range(n_img):
            cpu_reference_step(sd, spec, host_sets[0][k:k + 1], cores)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n_img / dt, "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": f"{n_img} image(s) of the first batch, restated reference on the host "
                                          f"(torch CPU fp32 + C polygon NMS), {dt:.1f} s"}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), file=out_stream, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: r101_b32 (configs[2]); r101_b16 (configs[3]) when WORLD_SIZE > 1")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--cpu-images", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-json", default="")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "r101_b16" if int(os.environ.get("WORLD_SIZE", "1")) > 1 or args.gpus > 1 else "r101_b32"
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
