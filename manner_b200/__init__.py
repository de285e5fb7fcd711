"""manner_b200: B200-native (sm_100a) scoring / ensemble / metrics hot path of MANNeR.

Importing the package is cheap and GPU-free; anything that computes goes through
libmanner_b200.so (``python -m manner_b200.build``) and raises if it is missing -- there is no CPU
or PyTorch fallback.
"""
from . import _native  # noqa: F401
from .data import Behaviours, from_segment_ids, synth_behaviours, synth_table, synth_workload  # noqa: F401

__all__ = ["Behaviours", "from_segment_ids", "synth_behaviours", "synth_table", "synth_workload", "ScoreEvaluator", "ops"]


def __getattr__(name):
    # ops / evaluator register torch custom ops on import; keep `import manner_b200` light
    import importlib

    if name in ("ops", "evaluator", "dist", "modules", "build"):
        return importlib.import_module("." + name, __name__)
    if name in ("ScoreEvaluator", "EvalResult", "DeviceBehaviours"):
        return getattr(importlib.import_module(".evaluator", __name__), name)
    raise AttributeError(name)
