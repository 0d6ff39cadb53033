"""padt_b200 -- B200-native (sm_100a) implementation of PaDT's generate/decode hot path.

Drop-in names of the reference package (``src/PaDT/__init__.py:1``):
``PaDTForConditionalGeneration``, ``VisonTextProcessingClass`` (sic), ``parseVRTintoCompletion``.
"""
from .config import PadtConfig, PRESETS, padt_pro_3b, padt_pro_7b, padt_tiny, padt_small
from .processor import VisonTextProcessingClass, parseVRTintoCompletion

__all__ = ["PaDTForConditionalGeneration", "VisonTextProcessingClass", "parseVRTintoCompletion", "PadtConfig", "PRESETS",
           "padt_pro_3b", "padt_pro_7b", "padt_tiny", "padt_small"]


def __getattr__(name):
    # the model class needs the CUDA library; keep `import padt_b200` (config / processor / host logic) usable without it
    if name == "PaDTForConditionalGeneration":
        from .modeling import PaDTForConditionalGeneration
        return PaDTForConditionalGeneration
    if name == "PadtEngine":
        from .engine import PadtEngine
        return PadtEngine
    raise AttributeError(name)
