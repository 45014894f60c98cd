"""YAML -> dataclass tree (API mirror of dll/configs/config_loader.py:15-204).

The reference spells out every field by hand; here one recursive builder walks the dataclass tree.  The places
where the reference loader's fallback differs from the dataclass default are kept in `_LOADER_DEFAULTS`.
"""
from dataclasses import dataclass, field, fields, is_dataclass
from pathlib import Path
from typing import Any, Dict

import yaml

from .base_config import DeviceConfig
from .model_config import ModelConfig
from .training_config import TrainingConfig


@dataclass
class PathsConfig:
    default_config: str = ""
    data_dir: str = ""
    output_dir: str = ""


@dataclass
class Config:
    paths: PathsConfig = field(default_factory=PathsConfig)
    device: DeviceConfig = field(default_factory=DeviceConfig)
    model: ModelConfig = field(default_factory=ModelConfig)
    training: TrainingConfig = field(default_factory=TrainingConfig)


# (dotted path) -> value used when the YAML omits the key (config_loader.py:65,87-88,101)
_LOADER_DEFAULTS = {
    "model.backbone.convert_to_grayscale": True,
    "model.keypoint_head.height": 56,
    "model.keypoint_head.width": 56,
    "model.heatmap_head.heatmap_size": (56, 56),
    "model.heatmap_head.deconv_kernel_sizes": [4, 4],
    "model.heatmap_head.deconv_channels": [256, 256],
}


def _build(cls, data: Dict[str, Any], path: str):
    data = data or {}
    kwargs = {}
    for f in fields(cls):
        p = f"{path}.{f.name}" if path else f.name
        sub = f.type if isinstance(f.type, type) else None
        if sub is not None and is_dataclass(sub):
            kwargs[f.name] = _build(sub, data.get(f.name, {}), p)
        elif f.name in data:
            v = data[f.name]
            kwargs[f.name] = tuple(v) if p == "model.heatmap_head.heatmap_size" else v
        elif p in _LOADER_DEFAULTS:
            kwargs[f.name] = _LOADER_DEFAULTS[p]
    return cls(**kwargs)


def load_config(config_path: str) -> Config:
    """Load configuration from a YAML file (same result object shape as the reference's load_config)."""
    config_path = Path(config_path)
    if not config_path.exists():
        raise FileNotFoundError(f"Config file not found: {config_path}")
    with open(config_path, 'r') as f:
        y = yaml.safe_load(f) or {}
    paths = _build(PathsConfig, y.get('paths'), "paths")
    device = _build(DeviceConfig, y.get('device'), "device")
    model_y = dict(y.get('model') or {})
    # ModelConfig.num_keypoints is taken from the keypoint_head section (config_loader.py:111)
    model_y['num_keypoints'] = (model_y.get('keypoint_head') or {}).get('num_keypoints', 17)
    model = _build(ModelConfig, model_y, "model")
    training = _build(TrainingConfig, y.get('training'), "training")
    training.device = device                      # the reference passes the top-level device section through
    return Config(paths=paths, device=device, model=model, training=training)
