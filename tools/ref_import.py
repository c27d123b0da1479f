"""Import the reference's own modules from /root/reference (build container only).

Used by tools/make_golden.py and by tests that re-validate the oracle when the reference is present.
The reference needs matplotlib/seaborn only for plots (metrics.py:13-14); they are absent in this image
and are stubbed.  Nothing is copied from the reference.
"""
import importlib.util
import sys
import types
from pathlib import Path

REF = Path("/root/reference")


def available() -> bool:
    return (REF / "pesquisa_v6" / "v6_pipeline" / "models.py").exists()


def _stub(name):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__dict__.setdefault("__path__", [])
        sys.modules[name] = m
    return sys.modules[name]


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns a namespace with the reference's models, pipeline, FGVC model, extraction and data hub."""
    for n in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        _stub(n)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    v6 = str(REF / "pesquisa_v6" / "v6_pipeline")
    if v6 not in sys.path:
        sys.path.insert(0, v6)
    ns = types.SimpleNamespace()
    ns.models = _load("ref_models", REF / "pesquisa_v6/v6_pipeline/models.py")
    ns.data_hub = _load("ref_data_hub", REF / "pesquisa_v6/v6_pipeline/data_hub.py")
    ns.pipe = _load("ref_pipe008", REF / "pesquisa_v6/scripts/008_run_pipeline_eval_v6.py")
    ns.fgvc = ns.pipe.fgvc_module
    ns.extract = _load("ref_extract005", REF / "pesquisa_v5/005_rearrange_video_YUV_420_10bit_LOSSLESS.py")
    return ns
