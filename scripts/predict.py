#!/usr/bin/env python
"""Keypoint prediction CLI on the B200-native `dll` drop-in (reference: scripts/predict.py:30-226).

    python scripts/predict.py --config configs/default_config.yaml --model ckpt.pth --input image_or_dir \
        [--gt label_file_or_dir] --output out_dir [--device cuda] [--precision bf16|fp32|fp32_strict] [--full-image-box | --detect]

Same command line, checkpoint handling and printout as the reference's script:
  * the checkpoint is either a training checkpoint {'epoch','model_state_dict','optimizer_state_dict','history'} (keys the
    model does not know are dropped) or a bare state_dict; both load with strict=False (predict.py:47-57);
  * person boxes come from a YOLO-pose label file, one person per line `class cx cy w h (x y v)*17`, normalised
    (predict.py:80-95, label format dll/data/dataloader.py:199-257);
  * the 17 COCO joints of the first person are printed as `(x, y)` in normalised image coordinates (predict.py:110-131).
The reference imports two modules that do not exist upstream (dll.data.create_dataloader and
dll.data.dataloader.visualize_keypoints_multi_person are importable only with matplotlib present) and writes no
output file although it logs one; here every image additionally gets `<stem>_keypoints.json` in --output.
Without a label file the reference passes bboxes=None and returns all-zero keypoints; --full-image-box substitutes the
box [0.5, 0.5, 1, 1] instead, --detect takes the boxes from the model's own person heads (the bare-tensor forward, csrc/kpd_detect.cu)
-- both new behaviour, off by default.
"""
import argparse
import json
import logging
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT / "keypoint-detection_b200"),):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import yaml

KEYPOINT_NAMES = [
    "nose", "left_eye", "right_eye", "left_ear", "right_ear", "left_shoulder", "right_shoulder", "left_elbow",
    "right_elbow", "left_wrist", "right_wrist", "left_hip", "right_hip", "left_knee", "right_knee", "left_ankle",
    "right_ankle"]                                                       # scripts/predict.py:110-115
IMAGE_SUFFIXES = ('.jpg', '.jpeg', '.png')


def load_config(config_path):
    with open(config_path, 'r') as f:
        return yaml.safe_load(f)


def build_model_config(config_dict):
    """predict.py:33-38: ModelConfig from the YAML's model section."""
    from dll.configs import ModelConfig, BackboneConfig, PersonDetectionConfig, KeypointHeadConfig
    m = config_dict['model']
    return ModelConfig(backbone=BackboneConfig(**m['backbone']), person_head=PersonDetectionConfig(**m['person_head']),
                       keypoint_head=KeypointHeadConfig(**m['keypoint_head']),
                       num_keypoints=m['keypoint_head']['num_keypoints'])


def extract_state_dict(checkpoint, model_keys):
    """predict.py:47-57: training checkpoint -> its 'model_state_dict' filtered to the keys the model has; a bare
    state_dict is returned unchanged."""
    if isinstance(checkpoint, dict) and 'model_state_dict' in checkpoint:
        known = set(model_keys)
        return {k: v for k, v in checkpoint['model_state_dict'].items() if k in known}
    return checkpoint


def load_model(model_path, config_dict, device, precision="bf16"):
    from dll.configs import TrainingConfig
    from dll.models import MultiPersonKeypointModel
    model = MultiPersonKeypointModel(build_model_config(config_dict), TrainingConfig(), precision=precision)
    try:
        checkpoint = torch.load(model_path, map_location="cpu")
    except Exception:                      # training checkpoints carry plain-Python history objects
        checkpoint = torch.load(model_path, map_location="cpu", weights_only=False)
    model.load_state_dict(extract_state_dict(checkpoint, model.state_dict().keys()), strict=False)
    model.to(device)
    model.eval()
    return model


def read_gt_boxes(gt_path):
    """predict.py:80-95: every line with >= 5 fields contributes one normalised (cx, cy, w, h) box -> Tensor[P,4] or None."""
    if gt_path is None or not Path(gt_path).exists():
        return None
    boxes = []
    with open(gt_path, 'r') as f:
        for line in f:
            data = line.strip().split()
            if len(data) >= 5:
                boxes.append([float(data[1]), float(data[2]), float(data[3]), float(data[4])])
    return torch.tensor(boxes, dtype=torch.float32).reshape(-1, 4) if boxes else None


def format_keypoints(keypoints):
    """predict.py:118-131: the printout for the first person.  keypoints: array-like [..., 17, 2] after squeeze()."""
    import numpy as np
    k = np.asarray(keypoints)
    lines = ["", "Predicted Keypoints:", "-" * 40, f"Keypoints shape: {k.shape}"]
    if k.ndim == 3:
        k = k[0]
    for i, name in enumerate(KEYPOINT_NAMES):
        if i < len(k):
            x, y = k[i]
            lines.append(f"{i + 1:2d}. {name:<15} ({x:.3f}, {y:.3f})")
    return lines


def predict_single_image(model, image_path, transform, device, gt_path=None, output_path=None, full_image_box=False, detect=False):
    from PIL import Image
    image = Image.open(image_path).convert('RGB')
    x = transform(image).unsqueeze(0)
    boxes = read_gt_boxes(gt_path)
    if boxes is None and full_image_box:
        boxes = torch.tensor([[0.5, 0.5, 1.0, 1.0]])
    batch = {'image': x, 'bboxes': boxes.unsqueeze(0) if boxes is not None else None}
    if boxes is None and detect:
        batch = {'image': x}                 # no 'bboxes' key: boxes from the person heads
    with torch.no_grad():
        outputs = model(batch)
    if boxes is None and detect:
        boxes = outputs['boxes'][0].cpu()
    keypoints = outputs['keypoints'].squeeze().cpu().numpy()
    for line in format_keypoints(keypoints):
        print(line)
    if output_path is not None:
        output_path = Path(output_path)
        output_path.parent.mkdir(parents=True, exist_ok=True)
        vis = outputs['visibilities'].squeeze().cpu().numpy()
        with open(output_path.with_name(output_path.stem + ".json"), 'w') as f:
            json.dump({"image": str(image_path), "keypoint_names": KEYPOINT_NAMES,
                       "boxes": boxes.tolist() if boxes is not None else [],
                       "keypoints": outputs['keypoints'].cpu().numpy().reshape(-1, len(KEYPOINT_NAMES), 2).tolist(),
                       "visibilities": vis.tolist()}, f)
    return keypoints[0] if keypoints.ndim == 3 else keypoints


def iter_inputs(input_path, gt_path):
    """(image, label-or-None) pairs: a single file, or every image of a directory with `<stem>.txt` from the --gt directory."""
    input_path = Path(input_path)
    gt_path = Path(gt_path) if gt_path else None
    if input_path.is_file():
        yield input_path, (gt_path if gt_path and gt_path.is_file() else None)
        return
    for img in sorted(input_path.glob('*.*')):
        if img.suffix.lower() in IMAGE_SUFFIXES:
            yield img, (gt_path / f"{img.stem}.txt" if gt_path else None)


def main(argv=None):
    parser = argparse.ArgumentParser(description='Predict keypoints in images (B200-native dll drop-in)')
    parser.add_argument('--config', type=str, required=True, help='Path to config file')
    parser.add_argument('--model', type=str, required=True, help='Path to model checkpoint')
    parser.add_argument('--input', type=str, required=True, help='Path to input image or directory')
    parser.add_argument('--gt', type=str, help='Path to ground truth label file or directory')
    parser.add_argument('--output', type=str, required=True, help='Path to output directory')
    parser.add_argument('--device', type=str, default='cuda', help='CUDA device (this build has no CPU path)')
    parser.add_argument('--precision', type=str, default='bf16', choices=['bf16', 'fp32', 'fp32_strict'])
    parser.add_argument('--full-image-box', action='store_true', help='no label file: use the box [0.5,0.5,1,1]')
    parser.add_argument('--detect', action='store_true', help='no label file: boxes from the person heads (new behaviour)')
    parser.add_argument('--host-transform', action='store_true', help="run ITransform with OpenCV on the host like the reference "
                        "(default: the bit-identical CUDA pipeline, kpd_itransform)")
    args = parser.parse_args(argv)

    logging.basicConfig(level=logging.INFO, format='%(asctime)s - %(levelname)s - %(message)s')
    config = load_config(args.config)
    device = torch.device(args.device)
    from dll.data.transforms import ITransform
    backbone = config['model']['backbone']
    transform = ITransform(img_size=backbone['input_size'], clip_limit=2.0, tile_size=(8, 8),
                           grayscale=backbone.get('in_channels', 3) == 1 or backbone.get('convert_to_grayscale', False),
                           device=None if args.host_transform else device)
    logging.info(f"Loading model from {args.model}")
    model = load_model(args.model, config, device, args.precision)
    n = 0
    for img, gt in iter_inputs(args.input, args.gt):
        out = Path(args.output) / f"{img.stem}_keypoints{img.suffix}"
        predict_single_image(model, img, transform, device, gt_path=gt, output_path=out, full_image_box=args.full_image_box, detect=args.detect)
        logging.info(f"Saved keypoints to {out.with_name(out.stem + '.json')}")
        n += 1
    logging.info(f"Prediction completed successfully! ({n} image(s))")
    return 0


if __name__ == '__main__':
    sys.exit(main())
