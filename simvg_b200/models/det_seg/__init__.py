from .base import BaseModel
from .one_stage import OneStageModel
from .mix_detr_mb import MIXDETRMB
