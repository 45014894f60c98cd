#!/usr/bin/env python
"""bench.py -- images/sec of forward + heat-map decode (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--precision bf16|fp32]

One "step" = one pass of the hot path (MultiPersonKeypointModel.predict: trunk -> top-k -> ROI-align -> head ->
decode) over one synthetic batch.  Workload at every N: BASELINE config C2 per GPU (B=256 images of 1x224x224,
one full-image box each, random-init weights of the reference architecture); ranks are independent replicas
(weak scaling, no data-path collective), timing = max over ranks of the CUDA-event time of K steps.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "keypoint-detection_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

METRIC = "images/sec (forward+heatmap decode)"
FLOP_PER_IMAGE = 9.42e9          # SURVEY.md 8d (S=224, P=1, dead FPN levels skipped)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "fp32_strict"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU")
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--persons", type=int, default=1)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]), burst=d["bf16_tflops"], source="measured")
    return dict(hbm=6650.0, tensor=1400.0, burst=1590.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# MobileNetV3-Small inverted-residual table (cin, expanded, cout, stride): the shapes the trunk kernels run on
MBV3 = [(16, 16, 16, 2), (16, 72, 24, 2), (24, 88, 24, 1), (24, 96, 40, 2), (40, 240, 40, 1), (40, 240, 40, 1),
        (40, 120, 48, 1), (48, 144, 48, 1), (48, 288, 96, 2), (96, 576, 96, 1), (96, 576, 96, 1)]


def trunk_rooflines(kern, B, S, pk, clocks):
    """Algorithmic bytes / flops of the trunk's kernel classes (layer input + output once per kernel) against the HBM copy
    peak and the nominal fp32 FMA peak (148 SMs x 128 lanes x 2 flop x the SM clock observed under load)."""
    out = {}
    h = S // 2
    dw_bytes = pw_flops = 0
    for cin, exp, cout, stride in MBV3:
        ho = h // stride
        dw_bytes += 4 * exp * (h * h + ho * ho)
        if exp != cin:
            pw_flops += 2 * h * h * cin * exp
        pw_flops += 2 * ho * ho * exp * cout
        h = ho
    pw_flops += 2 * h * h * 96 * 576
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    if "body.depthwise" in kern:
        out["depthwise (11 TMA-staged kernels; layer in+out bytes)"] = {
            "achieved_gbs": B * dw_bytes / (kern["body.depthwise"] * 1e-3) / 1e9, "peak_gbs": pk["hbm"]}
    if "body.pointwise" in kern:
        out["pointwise GEMMs (22 kernels: tcgen05 3xTF32, CUDA-core fp32 for K<=16 / narrow residual shapes; useful flops)"] = {
            "achieved_tflops": B * pw_flops / (kern["body.pointwise"] * 1e-3) / 1e12, "peak_tflops": fp32_peak,
            "peak_source": f"nominal fp32 FMA: 148 SM x 128 lanes x 2 x {sm_mhz:.0f} MHz"}
    if "fpn_laterals.lat0" in kern:
        px = B * (S // 2) ** 2
        out["lateral0 (1x1 16->128 + x4 top-down add; in+coarse+out bytes)"] = {
            "achieved_gbs": (px * (64 + 512) + B * (S // 8) ** 2 * 512) / (kern["fpn_laterals.lat0"] * 1e-3) / 1e9, "peak_gbs": pk["hbm"]}
    return out



def make_inputs(batch, size, persons):
    """Synthetic batch: randn images (the range Normalize(0.5, 0.5) produces), one full-image box per image for
    P=1 (the canonical throughput point) or seeded random boxes for P>1."""
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, 1, size, size, generator=g)
    if persons == 1:
        boxes = torch.tensor([0.5, 0.5, 1.0, 1.0]).repeat(batch, 1, 1)
    else:
        c = 0.25 + 0.5 * torch.rand(batch, persons, 2, generator=g)
        wh = 0.2 + 0.5 * torch.rand(batch, persons, 2, generator=g)
        boxes = torch.cat([c, wh], dim=-1)
    return x.contiguous(), boxes.contiguous()


def cpu_reference(args, seconds, steps=None, warmup=1):
    """The reference's CPU forward restated by oracle/ (kind 'port': the reference is pure Python and cannot travel to
    the GPU box), all host threads, config C1 sample: B=8 images per pass."""
    from oracle import model as OM
    from oracle.weights import synth_state_dict
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth_state_dict(0)
    x, boxes = make_inputs(8, args.size, args.persons)
    for _ in range(warmup):
        OM.forward(sd, x, boxes)
    times = []
    t_end = time.perf_counter() + seconds
    while (steps is None and time.perf_counter() < t_end) or (steps is not None and len(times) < steps):
        t0 = time.perf_counter(); OM.forward(sd, x, boxes); times.append(time.perf_counter() - t0)
        if steps is None and len(times) >= 200:
            break
    total = float(np.sum(times))
    return {"value": 8 * len(times) / total, "unit": "images/sec", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{len(times)} passes of B=8 x 1x{args.size}x{args.size}, P={args.persons} (config C1) through oracle.model.forward "
                      f"(reference-structured per-box loop, torch CPU fp32); best pass {8 / min(times):.1f} img/s",
            "ms_per_pass": 1e3 * total / len(times)}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    steps = max(1, args.steps)
    cb = cpu_reference(args, 0, steps=steps, warmup=max(1, min(args.warmup, 2)))
    line = {"metric": METRIC, "value": cb["value"], "unit": "images/sec", "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_pass"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"C2: B={args.batch} per GPU, 1x{args.size}x{args.size} images, P={args.persons} box/image, forward+decode, "
                                   f"random-init weights (seed 0)",
                       "global_batch": args.batch * world, "parallelism": f"dp{world} replicas", "precision": "fp32 (CPU)",
                       "sample": f"bounded sample of that workload: each step is B=8 images (1x{args.size}x{args.size}, P={args.persons}) "
                                 f"through the CPU port of the reference on all host cores; rank 0 only", "images_per_step": 8},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)

    from dll import MultiPersonKeypointModel, _native
    from dll.configs.config_loader import load_config

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
    torch.manual_seed(0)                 # random-init weights of the reference architecture (no checkpoint exists upstream)
    model = MultiPersonKeypointModel(cfg.model, cfg.training, precision=args.precision)
    model = model.to(dev).eval()
    lib = _native.load()

    B, S, P = args.batch, args.size, args.persons
    x_host, b_host = make_inputs(B, S, P)
    x_pin, b_pin = x_host.pin_memory(), b_host.pin_memory()
    x_dev, b_dev = x_pin.to(dev), b_pin.to(dev)
    batch_dev = {"image": x_dev, "bboxes": b_dev}

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        return ms

    def step_resident():
        model.predict(batch_dev)

    kp_host = torch.empty(B, P, 1, 17, 2).pin_memory(); vis_host = torch.empty(B, P, 1, 17, 3).pin_memory()

    def step_e2e():
        out = model.predict({"image": x_pin, "bboxes": b_pin})          # H2D of this step's inputs from pinned memory
        kp_host.copy_(out["keypoints"], non_blocking=True); vis_host.copy_(out["visibilities"], non_blocking=True)
        torch.cuda.current_stream().synchronize()                        # the caller holds the step's result

    for _ in range(max(3, args.warmup)):
        step_resident()
    lib.kpd_launch_count(1)
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local)
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, args.steps)
    launches = int(lib.kpd_launch_count(1))
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # per-kernel CUDA-event timing (library-side events on the launching stream), outside the timed region
    lib.kpd_profile_enable(1)
    prof_steps = 5
    for _ in range(prof_steps):
        step_resident()
    torch.cuda.synchronize()
    import ctypes as C
    kern = {}
    for tag in ("body", "fpn_laterals", "fpn0_conv", "pool", "scores_topk", "roi_align", "head_attention", "head_conv1",
                "head_conv2", "head_conv3", "head_final", "decode"):
        t, n = C.c_double(0), C.c_int64(0)
        if lib.kpd_profile_read(tag.encode(), C.byref(t), C.byref(n)) == 0 and n.value:
            kern[tag] = t.value / n.value
    # kernel classes inside the trunk scopes: ms per step summed over the class's launches
    for tag in ("body.stem", "body.pointwise", "body.depthwise", "body.se", "fpn_laterals.lat0"):
        t, n = C.c_double(0), C.c_int64(0)
        if lib.kpd_profile_read(tag.encode(), C.byref(t), C.byref(n)) == 0 and n.value:
            kern[tag] = t.value / prof_steps
    lib.kpd_profile_enable(0)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    pk = peaks()
    ms_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    n_roi = B * P
    conv2_flops = 2.0 * n_roi * 56 * 56 * 256 * 2304
    dom = "head_conv2"
    roof = None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")     # dram bytes/launch from the committed ncu capture
    if os.path.exists(tpath):
        t = json.load(open(tpath)).get(f"{dom}_{args.precision}_B{B}")
        traffic = t["dram_bytes_per_launch"] if t else None
    if dom in kern:
        ach = conv2_flops / (kern[dom] * 1e-3) / 1e12
        peak = pk["tensor"] if args.precision == "bf16" else pk["tensor"] / 2
        roof = {"bound": "tensor", "kernel": "conv_umma_kernel (head conv 256->256, 3x3)", "achieved": ach, "peak": peak,
                "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": f"{pk['source']} bf16 sustained" + ("" if args.precision == "bf16" else " / 2 (tf32 nominal half rate)"),
                "flops_per_launch": conv2_flops, "ms_per_launch": kern[dom]}
    line = {"metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 head convs, tf32 FPN-0, 3xTF32 / fp32 trunk, fp32 accumulate" if args.precision == "bf16" else ("tf32 convs, 3xTF32 / fp32 trunk, fp32 accumulate" if args.precision == "fp32" else "fp32"),
            "data": "synthetic", "impl": "native",
            "config": {"workload": f"C2: B={B} per GPU, 1x{S}x{S} images, P={P} box/image, forward+decode, random-init weights (seed 0)",
                       "global_batch": B * world, "parallelism": f"dp{world} replicas", "precision": args.precision,
                       "l2": "per-step working set (workspace >1 GB) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "images/sec", "h2d_bytes_per_step": int(x_pin.numel() * 4 + b_pin.numel() * 4),
                    "d2h_bytes_per_step": int(kp_host.numel() * 4 + vis_host.numel() * 4), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof,
            "kernels_ms": kern,
            "roofline_other": {
                "fpn0_conv (tcgen05 tf32; reference-equivalent 3.70 GFLOP/img, executed 54/144 of it in composed form)": None if "fpn0_conv" not in kern else
                    {"reference_equivalent_tflops": 2.0 * B * (S // 2) ** 2 * 128 * 1152 / (kern["fpn0_conv"] * 1e-3) / 1e12,
                     "achieved_tflops": (54.0 / 144.0) * 2.0 * B * (S // 2) ** 2 * 128 * 1152 / (kern["fpn0_conv"] * 1e-3) / 1e12,
                     "peak_tflops": pk["tensor"] / 2,
                     "note": "composed form (kpd_fpn0.cu): conv3x3(L0 f0) with K=144 plus 36 coarse-neighbour matrices instead of 16x9 taps"},
                "body (13 fused kernels, fp32 FMA; algorithmic bytes 0.77 M elem x 4 B per image)": None if "body" not in kern else
                    {"achieved_gbs": B * 0.77e6 * 4 / (kern["body"] * 1e-3) / 1e9, "peak_gbs": pk["hbm"]},
                "decode (fp32 heat maps 17x56x56 per ROI)": None if "decode" not in kern else
                    {"achieved_gbs": n_roi * (17 * 3136 * 4 + 340) / (kern["decode"] * 1e-3) / 1e9, "peak_gbs": pk["hbm"]},
                "roi_align (64 selected channels of the ROI's FPN-0 window in, 56x56x64 fp32 out)": None if "roi_align" not in kern else
                    {"achieved_gbs": n_roi * (64 * (S // 2) ** 2 * 4 * (1.0 if P == 1 else 0.25) + 3136 * 64 * 4) / (kern["roi_align"] * 1e-3) / 1e9,
                     "peak_gbs": pk["hbm"], "note": "P=1: the box is the whole map; P>1: seeded boxes cover ~1/4 of the map on average"},
                "head_attention (56x56x64 fp32 in, head-precision out)": None if "head_attention" not in kern else
                    {"achieved_gbs": n_roi * 3136 * 64 * (4 + (2 if args.precision == "bf16" else 4)) / (kern["head_attention"] * 1e-3) / 1e9,
                     "peak_gbs": pk["hbm"]},
                **trunk_rooflines(kern, B, S, pk, clocks)},
            "model_tflops": value * FLOP_PER_IMAGE / 1e12 / world}
    if not args.no_cpu_baseline and world == 1:          # rank 0 at N=1 only (task contract)
        line["cpu_baseline"] = cpu_reference(args, args.cpu_baseline_seconds)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
