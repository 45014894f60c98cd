from .base_config import BaseConfig, DeviceConfig
from .model_config import (ModelConfig, BackboneConfig, PersonDetectionConfig, KeypointHeadConfig,
                           HeatmapHeadConfig)
from .training_config import (TrainingConfig, OptimizerConfig, AugmentationConfig, LossConfig,
                              LRSchedulerConfig, WeightedLossConfig)

__all__ = ['BaseConfig', 'DeviceConfig', 'ModelConfig', 'BackboneConfig', 'PersonDetectionConfig',
           'KeypointHeadConfig', 'HeatmapHeadConfig', 'TrainingConfig', 'OptimizerConfig',
           'AugmentationConfig', 'LossConfig', 'LRSchedulerConfig']
