"""Person-box selection of the bare-tensor forward path -- TEST INFRASTRUCTURE (checker only, never imported by the product).

NEW BEHAVIOUR, not a restatement of the reference: upstream's bare-tensor path returns raw head outputs and crashes
(dll/models/keypoint_model.py:114-119, dll/models/person_head.py:141-166), so there is nothing to pin against.  This numpy
function states the semantics that keypoint-detection_b200/csrc/kpd_detect.cu defines, built only from the reference's own
ingredients -- the anchor order of person_head.py:45-51, PersonDetectionConfig.conf_threshold / nms_iou_threshold and the
reference's greedy NMS (person_head.py:96-139, restated in oracle.decode) -- so that the CUDA path has an independent check.
"""
import numpy as np

from .decode import non_max_suppression as _nms


def detect_persons(box, cls, height, width, size, anchor_sizes=(32, 64, 128), conf_threshold=0.3, iou_threshold=0.3,
                   top_m=100, max_persons=4):
    """box [B,H*W,36], cls [B,H*W,9] (level-0 head outputs) -> (boxes [B,max_persons,4] cxcywh zero-padded, count [B])."""
    box = np.asarray(box, np.float32); cls = np.asarray(cls, np.float32)
    B = box.shape[0]
    out = np.zeros((B, max_persons, 4), np.float32)
    count = np.zeros(B, np.int32)
    ratios = (0.5, 1.0, 2.0)
    for b in range(B):
        logit = cls[b].reshape(-1)
        # top_m by logit (sigmoid is monotone), ties towards the lower candidate index
        order = np.lexsort((np.arange(logit.size), -logit.astype(np.float64)))[:top_m]
        cand = np.zeros((len(order), 4), np.float32); score = np.zeros(len(order), np.float32)
        for r, c in enumerate(order):
            pixel, an = divmod(int(c), 9)
            y, x = divmod(pixel, width)
            s = np.float32(anchor_sizes[an // 3]); ratio = np.float32(ratios[an % 3])
            aw = np.float32(s * ratio) * np.float32(1.0 / size); ah = np.float32(s / ratio) * np.float32(1.0 / size)
            acx = np.float32((x + 0.5)) / np.float32(width); acy = np.float32((y + 0.5)) / np.float32(height)
            d = box[b, pixel, an * 4:an * 4 + 4]
            cx = np.clip(acx + d[0] * aw, 0, 1); cy = np.clip(acy + d[1] * ah, 0, 1)
            w = min(aw * np.exp(np.clip(d[2], -4, 4)), 1.0); h = min(ah * np.exp(np.clip(d[3], -4, 4)), 1.0)
            cand[r] = (cx, cy, w, h)
            score[r] = 1.0 / (1.0 + np.exp(-np.float64(logit[c])))
        nv = int((score >= np.float32(conf_threshold)).sum())
        keep = _nms(cand, score, iou_threshold, max_persons)
        kept = [int(k) for k in keep if k < nv]
        for p, k in enumerate(kept[:max_persons]):
            out[b, p] = cand[k]
        count[b] = min(len(kept), max_persons)
    return out, count
