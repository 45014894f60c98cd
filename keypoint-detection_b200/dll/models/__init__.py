from .keypoint_model import MultiPersonKeypointModel
from .backbone import BACKBONE, MobileNetV3Wrapper, LightweightFPN
from .person_head import PERSON_HEAD
from .keypoint_head import KEYPOINT_HEAD
from .heatmap_head import HeatmapHead

__all__ = ['MultiPersonKeypointModel', 'BACKBONE', 'PERSON_HEAD', 'KEYPOINT_HEAD', 'MobileNetV3Wrapper', 'HeatmapHead', 'LightweightFPN']
