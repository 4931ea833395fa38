"""oracle/ref_shims.py -- TEST INFRASTRUCTURE ONLY.  Makes the UNMODIFIED reference Python importable in the build
container (where /root/reference exists) so its own functions can generate golden vectors.  The reference imports
onnxruntime / tensorrt / pycuda / lap at module import and uses numpy aliases removed in numpy >= 1.24
(SURVEY.md Appendix B); the shims stub the absent modules, restore `np.float`, restate `lap.lapjv` on scipy and
guard one empty-array comparison.  /root/reference does not exist on the GPU box: nothing under tests -m gpu,
smoke() or bench.py imports this module.
"""
import importlib.machinery
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ObjectDetector"))


def install():
    assert available(), "reference tree not present"
    for n in ("onnxruntime", "tensorrt", "pycuda", "pycuda.driver"):
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__spec__ = importlib.machinery.ModuleSpec(n, None)     # torch._dynamo probes find_spec() on loaded modules
            sys.modules[n] = m
    sys.modules["pycuda"].driver = sys.modules["pycuda.driver"]
    if not hasattr(np, "float"):
        np.float = float
    lap = types.ModuleType("lap")
    lap.__spec__ = importlib.machinery.ModuleSpec("lap", None)

    def lapjv(cost, extend_cost=False, cost_limit=np.inf, return_cost=True):
        from scipy.optimize import linear_sum_assignment
        cost = np.asarray(cost, float)
        nr, nc = cost.shape
        n = nr + nc
        ext = np.full((n, n), cost_limit / 2.0)
        ext[nr:, nc:] = 0
        ext[:nr, :nc] = cost
        r, c = linear_sum_assignment(ext)
        x = np.full(n, -1)
        y = np.full(n, -1)
        x[r] = c
        y[c] = r
        x[x >= nc] = -1
        y[y >= nr] = -1
        x, y = x[:nr], y[:nc]
        return cost[np.nonzero(x != -1)[0], x[x != -1]].sum(), x, y

    lap.lapjv = lapjv
    sys.modules["lap"] = lap
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from ObjectDetector.utils import Scaler
    if not getattr(Scaler, "_b200_guarded", False):
        orig = Scaler.convert_kpss_coordinate
        Scaler.convert_kpss_coordinate = lambda self, k: (np.array(k) if np.array(k).size == 0 else orig(self, np.array(k)))
        Scaler._b200_guarded = True


class FakeEngine:
    """Stands in for OnnxEngine: returns canned output tensors (the detectors only use this protocol)."""

    def __init__(self, in_shape, out_shapes, out_names, outputs_fn):
        self.framework_type = "fake"
        self.providers = "fake"
        self.engine_dtype = np.float32
        self._in, self._out, self._names, self._fn = in_shape, out_shapes, out_names, outputs_fn
        self.last_input = None

    def get_engine_input_shape(self):
        return self._in

    def get_engine_output_shape(self):
        return self._out, self._names

    def engine_inference(self, x):
        self.last_input = x
        return self._fn(x)
