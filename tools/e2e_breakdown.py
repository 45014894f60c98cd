"""Where the end-to-end (host-input) step spends its time: H2D alone, a synchronised resident step, the full e2e step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "keypoint-detection_b200"), ROOT]
import torch
from dll import MultiPersonKeypointModel
from dll.configs.config_loader import load_config
cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
m = MultiPersonKeypointModel(cfg.model, cfg.training).cuda().eval()
B = 256
x = torch.randn(B, 1, 224, 224).pin_memory(); bx = torch.tensor([0.5, 0.5, 1.0, 1.0]).repeat(B, 1, 1).pin_memory()
xd, bd = x.cuda(), bx.cuda()
kp = torch.empty(B, 1, 1, 17, 2).pin_memory()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def h2d(): x.to("cuda", non_blocking=True); torch.cuda.synchronize()
def res_sync(): m.predict({"image": xd, "bboxes": bd}); torch.cuda.synchronize()
def res_nosync(): m.predict({"image": xd, "bboxes": bd})
def e2e():
    o = m.predict({"image": x, "bboxes": bx}); kp.copy_(o["keypoints"], non_blocking=True); torch.cuda.synchronize()
def host_only():
    t0 = time.perf_counter(); m.predict({"image": xd, "bboxes": bd}); return time.perf_counter() - t0
print("h2d 51 MB + sync      ms", round(t(h2d), 3))
print("resident, no sync     ms", round(t(res_nosync), 3))
print("resident, sync / step ms", round(t(res_sync), 3))
print("e2e                   ms", round(t(e2e), 3))
torch.cuda.synchronize(); print("host time of one predict() call (enqueue only) ms", round(min(host_only() for _ in range(10)) * 1e3, 3))

# host-side cost of the C-ABI call itself vs the Python wrapper around it
from dll import _native as N
import ctypes as C
plan, dev = m._plan()
lib = N.load()
P, S = 1, 224
prec = N.PRECISIONS[m.precision]
L = plan.layout(B, S, B * P, prec)
ws = torch.empty(L.total_bytes + 1024, dtype=torch.uint8, device="cuda")
heat = torch.empty(B, P, 17, 56, 56, device="cuda"); kpts = torch.empty(B, P, 1, 17, 2, device="cuda"); vis = torch.empty(B, P, 1, 17, 3, device="cuda")
def call():
    N.check(lib.kpd_forward(plan.handle, prec, N.ptr(xd), B, S, N.ptr(bd), P, None, N.ptr(heat), N.ptr(kpts), N.ptr(vis), N.ptr(ws), L.total_bytes, N.stream_ptr(dev)), "kpd_forward")
for _ in range(3): call()
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    torch.cuda.synchronize(); t0 = time.perf_counter(); call(); best = min(best, time.perf_counter() - t0)
print("host time of the bare kpd_forward C call ms", round(best * 1e3, 3))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): plan.layout(B, S, B * P, prec)
print("plan.layout ms", round((time.perf_counter() - t0) / 10 * 1e3, 4))
