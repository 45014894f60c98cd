"""CPU restatement of the standalone heatmap decoders and person NMS (test infrastructure; numpy).

  decode_heatmaps               dll/models/heatmap_head.py:265-296
  decode_heatmaps_subpixel      dll/models/heatmap_head.py:298-370 (window_size=3 default)
  decode_heatmaps_soft_argmax   dll/models/heatmap_head.py:372-413
  model_decode                  dll/models/keypoint_model.py:250-313 (soft-argmax + 3-class visibility)
  box_iou / non_max_suppression dll/models/person_head.py:54-139

Arithmetic is float32 like the reference.  argmax is "first maximum wins" (torch.max on CPU).
"""
import numpy as np

f32 = np.float32


def _flat(h):
    h = np.asarray(h, dtype=np.float32)
    if h.ndim == 3:
        h = h[None]
    B, K, H, W = h.shape
    return h, h.reshape(B, K, H * W)


def decode_heatmaps(heatmaps):
    """-> (keypoints [B,K,2] = (x/(W-1), y/(H-1)), scores [B,K], argmax int64 [B,K])."""
    h, flat = _flat(heatmaps)
    B, K, H, W = h.shape
    idx = flat.argmax(axis=2)
    mx = np.take_along_axis(flat, idx[..., None], axis=2)[..., 0]
    kp = np.stack([(idx % W).astype(np.float32) / f32(W - 1), (idx // W).astype(np.float32) / f32(H - 1)], axis=-1)
    return kp.astype(np.float32), mx, idx.astype(np.int64)


def decode_heatmaps_subpixel(heatmaps, window_size=3):
    """3x3 (border-clipped) centre of mass around the argmax; zero-mass windows give (0,0,0)."""
    h, flat = _flat(heatmaps)
    B, K, H, W = h.shape
    idx = flat.argmax(axis=2)
    xr, yr = idx % W, idx // W
    pad = window_size // 2
    kp = np.zeros((B, K, 2), dtype=np.float32)
    sc = np.zeros((B, K), dtype=np.float32)
    for b in range(B):
        for k in range(K):
            xc, yc = int(xr[b, k]), int(yr[b, k])
            xs, xe = max(0, xc - pad), min(W, xc + pad + 1)
            ys, ye = max(0, yc - pad), min(H, yc + pad + 1)
            win = h[b, k, ys:ye, xs:xe]
            tot = win.sum(dtype=np.float32)
            if tot > 0:
                xg = np.arange(xs, xe, dtype=np.float32)[None, :]
                yg = np.arange(ys, ye, dtype=np.float32)[:, None]
                kp[b, k, 0] = (xg * win).sum(dtype=np.float32) / tot
                kp[b, k, 1] = (yg * win).sum(dtype=np.float32) / tot
                sc[b, k] = win.max()
    kp[..., 0] /= f32(W - 1)
    kp[..., 1] /= f32(H - 1)
    return kp, sc, idx.astype(np.int64)


def _softmax_xy(flat, H, W, temperature=1.0):
    z = flat.astype(np.float32) / f32(temperature)
    z = z - z.max(axis=2, keepdims=True)
    e = np.exp(z, dtype=np.float32)
    p = e / e.sum(axis=2, keepdims=True, dtype=np.float32)
    B, K = p.shape[:2]
    p = p.reshape(B, K, H, W)
    x = (p * np.arange(W, dtype=np.float32)[None, None, None, :]).sum(axis=(2, 3), dtype=np.float32) / f32(W - 1)
    y = (p * np.arange(H, dtype=np.float32)[None, None, :, None]).sum(axis=(2, 3), dtype=np.float32) / f32(H - 1)
    return np.stack([x, y], axis=-1).astype(np.float32)


def decode_heatmaps_soft_argmax(heatmaps, temperature=1.0):
    h, flat = _flat(heatmaps)
    B, K, H, W = h.shape
    return _softmax_xy(flat, H, W, temperature), flat.max(axis=2)


def model_decode(heatmaps):
    """MultiPersonKeypointModel.decode_heatmap: (kp [B,K,2], vis one-hot [B,K,3], conf [B,K])."""
    h, flat = _flat(heatmaps)
    B, K, H, W = h.shape
    kp = _softmax_xy(flat, H, W, 1.0)
    mx = flat.max(axis=2)
    conf = (f32(1.0) / (f32(1.0) + np.exp(-mx, dtype=np.float32))).astype(np.float32)
    cls = np.where(conf < f32(0.3), 0, np.where(conf < f32(0.7), 1, 2))
    vis = np.eye(3, dtype=np.float32)[cls]
    return kp, vis, conf


def box_iou(b1, b2):
    """cxcywh IoU matrix [N,M]; union + 1e-16 (person_head.py:54-94)."""
    b1 = np.asarray(b1, dtype=np.float32); b2 = np.asarray(b2, dtype=np.float32)
    h = f32(2)
    a = [b1[:, 0] - b1[:, 2] / h, b1[:, 1] - b1[:, 3] / h, b1[:, 0] + b1[:, 2] / h, b1[:, 1] + b1[:, 3] / h]
    b = [b2[:, 0] - b2[:, 2] / h, b2[:, 1] - b2[:, 3] / h, b2[:, 0] + b2[:, 2] / h, b2[:, 1] + b2[:, 3] / h]
    ix1 = np.maximum(a[0][:, None], b[0]); iy1 = np.maximum(a[1][:, None], b[1])
    ix2 = np.minimum(a[2][:, None], b[2]); iy2 = np.minimum(a[3][:, None], b[3])
    inter = np.clip(ix2 - ix1, 0, None) * np.clip(iy2 - iy1, 0, None)
    a1 = (a[2] - a[0]) * (a[3] - a[1]); a2 = (b[2] - b[0]) * (b[3] - b[1])
    union = a1[:, None] + a2 - inter
    return (inter / (union + f32(1e-16))).astype(np.float32)


def non_max_suppression(boxes, scores, iou_threshold=0.2, max_output_size=None):
    """Greedy NMS, keep order = descending score (person_head.py:96-139). Returns int64 indices.

    Ties in `scores` are implementation-defined in the reference (unstable sort); this restatement
    breaks them towards the lower index, and tests use tie-free scores.
    """
    boxes = np.asarray(boxes, dtype=np.float32); scores = np.asarray(scores, dtype=np.float32)
    order = np.argsort(-scores, kind="stable")
    keep = []
    while order.size > 0:
        if order.size == 1:
            keep.append(int(order[0])); break
        i = int(order[0]); keep.append(i)
        if max_output_size and len(keep) >= max_output_size:
            break
        order = order[1:]
        iou = box_iou(boxes[i:i + 1], boxes[order])[0]
        order = order[iou <= f32(iou_threshold)]
    return np.asarray(keep, dtype=np.int64)
