import os, sys
ROOT="/root/repo"; sys.path[:0]=[ROOT+"/keypoint-detection_b200", ROOT, ROOT+"/tests"]
import numpy as np, torch
from dll import MultiPersonKeypointModel
from dll.configs.config_loader import load_config
from oracle import model as OM
cfg = load_config(ROOT+"/configs/default_config.yaml")
m = MultiPersonKeypointModel(cfg.model, cfg.training).cuda().eval()
g = np.load(ROOT+"/tests/golden/roi_align.npz")
feat = torch.from_numpy(g["feat"])
for i, roi in enumerate(g["rois"]):
    b = int(roi[0]); x1, y1, x2, y2 = [float(v) / 64.0 for v in roi[1:]]
    box = torch.tensor([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], dtype=torch.float32)
    corners = torch.stack([box[0] - box[2] / 2, box[1] - box[3] / 2, box[0] + box[2] / 2, box[1] + box[3] / 2]).clamp(0, 1) * 64
    want = OM.roi_align_legacy(feat[b], corners.view(1, 4).numpy())[0].numpy()
    got = m.extract_roi_features(feat[b:b + 1], box)[0].cpu().numpy()
    d = np.abs(got-want)
    print(i, roi, corners.numpy(), "max diff", d.max(), "at", np.unravel_index(d.argmax(), d.shape), "vs golden", np.abs(got-g["out"][i]).max())
