"""`dll` -- B200-native drop-in for the inference hot path of Dane-ME/keypoint-detection.

Public surface mirrors the reference's dll/__init__.py:1-16 for the hot path (model + configs); data loading and
training are out of scope of this build.
"""
from .models import MultiPersonKeypointModel
from .configs import ModelConfig, TrainingConfig

__version__ = '1.0.0'

__all__ = ['MultiPersonKeypointModel', 'ModelConfig', 'TrainingConfig']
