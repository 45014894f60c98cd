"""Base dataclass config (API mirror of the reference's dll/configs/base_config.py:11-64)."""
import os
from dataclasses import dataclass, fields, is_dataclass
from pathlib import Path
from typing import Dict

import yaml


@dataclass
class DeviceConfig:
    """dll/configs/base_config.py:11-17"""
    type: str = 'auto'
    force_cpu: bool = False
    mixed_precision: bool = True
    pin_memory: bool = True


@dataclass
class BaseConfig:
    @classmethod
    def get_config_path(cls) -> str:
        """$DLL_CONFIG_PATH, else <project root>/configs/default_config.yaml (base_config.py:21-32)."""
        env = os.getenv('DLL_CONFIG_PATH')
        if env:
            return env
        here = Path(__file__).resolve()
        for parent in here.parents:
            cand = parent / 'configs' / 'default_config.yaml'
            if cand.exists() and parent.name != 'dll':
                return str(cand)
        return str(here.parents[3] / 'configs' / 'default_config.yaml')

    def to_dict(self) -> Dict:
        return {f.name: getattr(self, f.name) for f in fields(self)}

    @classmethod
    def from_dict(cls, config_dict: Dict) -> 'BaseConfig':
        return cls(**config_dict)

    @classmethod
    def from_default(cls, key: str = None) -> 'BaseConfig':
        with open(cls.get_config_path()) as f:
            data = yaml.safe_load(f)
        if key:
            for part in key.split('.'):
                data = data[part]
        return cls.from_dict(data)

    def save_yaml(self, yaml_path: str) -> None:
        def plain(v):
            if is_dataclass(v):
                return {f.name: plain(getattr(v, f.name)) for f in fields(v)}
            if isinstance(v, tuple):
                return list(v)
            return v
        with open(yaml_path, 'w') as f:
            yaml.dump(plain(self), f)

    def validate(self) -> None:
        pass
