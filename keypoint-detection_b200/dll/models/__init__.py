from .keypoint_model import MultiPersonKeypointModel
from .backbone import MobileNetV3Wrapper, LightweightFPN
from .person_head import PERSON_HEAD
from .heatmap_head import HeatmapHead

__all__ = ['MultiPersonKeypointModel', 'PERSON_HEAD', 'MobileNetV3Wrapper', 'HeatmapHead', 'LightweightFPN']
