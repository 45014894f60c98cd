"""cProfile of the Python side of predict() (device-resident inputs): where the host time of a forward goes."""
import os, sys, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "keypoint-detection_b200"), ROOT]
import torch
from dll import MultiPersonKeypointModel
from dll.configs.config_loader import load_config
cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
m = MultiPersonKeypointModel(cfg.model, cfg.training).cuda().eval()
B = 256
xd = torch.randn(B, 1, 224, 224, device="cuda"); bd = torch.tensor([0.5, 0.5, 1.0, 1.0], device="cuda").repeat(B, 1, 1)
for _ in range(3): m.predict({"image": xd, "bboxes": bd})
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(50): m.predict({"image": xd, "bboxes": bd})
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
