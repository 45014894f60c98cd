"""Time the dense-conv kernel variants (kpd_dense_conv) at BASELINE config C2 sizes."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "keypoint-detection_b200"), ROOT]
import torch
from dll import MultiPersonKeypointModel, _native as N
from dll.configs.config_loader import load_config
cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
m = MultiPersonKeypointModel(cfg.model, cfg.training).cuda().eval()
plan, dev = m._plan()
lib = N.load()
res = {}
for conv, n, hw, cin, cout in [(0, 256, 112, 128, 128), (1, 256, 56, 64, 256), (2, 256, 56, 256, 256), (3, 256, 56, 256, 64)]:
    flops = 2.0 * n * hw * hw * cout * cin * 9
    x32 = torch.randn(n, hw, hw, cin, device="cuda")
    for impl, name, inp_dtype in [(1, "tf32", torch.float32), (3, "tf32_halo", torch.float32), (2, "bf16", torch.bfloat16), (4, "bf16_halo", torch.bfloat16)]:
        if conv == 0 and impl in (2, 4) and False:
            continue
        x = x32.to(inp_dtype)
        out = torch.empty(n, hw, hw, cout, device="cuda", dtype=inp_dtype)
        def run():
            N.check(lib.kpd_dense_conv(plan.handle, conv, impl, N.ptr(x), N.ptr(out), n, hw, hw, N.stream_ptr(dev)), "conv")
        for _ in range(3): run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): run()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        res[f"conv{conv}_{name}"] = {"ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1)}
        print(f"conv{conv} {cin}->{cout} @{hw} {name:10s} {ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s", flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_conv.json"), "w"), indent=1)
