"""simvg_b200.models mirrors the reference's `simvg.models` plugin surface (simvg/models/__init__.py:1-8): the five
registries plus the three classes SimVG's 53 configs actually resolve (BEIT3, TextGuidedQuerySelectKDDETRHead, MIXDETRMB)."""
from .builder import (FUSIONS, HEADS, LAN_ENCODERS, MODELS, VIS_ENCODERS, build_fusion, build_head, build_lan_enc,
                      build_model, build_vis_enc)
from .det_seg import *  # noqa: F401,F403
from .heads import *  # noqa: F401,F403
from .vis_encs import *  # noqa: F401,F403
from .utils import ExponentialMovingAverage  # noqa: F401,E402
