"""Import the UNMODIFIED CODD reference files in this container.

TEST / BASELINE INFRASTRUCTURE ONLY.  Used by ``oracle/gen_golden.py`` (fixture generation), by
the ``-m "not gpu"`` tests that pin the oracle restatement against the reference, and by the
baseline legs of ``bench.py`` (``--impl reference``, ``cpu_baseline``, ``gpu_eager_baseline``), which time
the UNMODIFIED reference modules from the git-ignored copy under ``baseline/_ref`` that
``__graft_entry__.build()`` makes (the snapshot `gpurun` ships carries it to the GPU box).  Nothing on the
product path imports this module.

The reference needs mmcv / mmseg (registry, BaseModule, init helpers) and imports
pytorch3d / lietorch at module scope.  None is installable here, so ``oracle/_shim``
provides import-compatible stand-ins (registry + no-op decorators; the pytorch3d /
lietorch stubs raise if they are ever *called*).  ``model`` is registered as a
namespace package so that ``model/__init__.py`` (which imports every sub-package)
is not executed; the stereo / fusion / utils files themselves run unmodified.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "_shim")


def reference_root():
    """$CODD_REF -> baseline/_ref (the git-ignored copy __graft_entry__.build() makes so that the unmodified reference
    travels to the GPU box with the snapshot) -> /root/reference (build container)."""
    cands = [os.environ.get("CODD_REF"), os.path.join(os.path.dirname(_HERE), "baseline", "_ref"), "/root/reference"]
    for root in cands:
        if root and os.path.isdir(os.path.join(root, "model", "stereo")):
            return root
    return None


def available():
    return reference_root() is not None


_loaded = {}


def load():
    """Returns a namespace with the reference modules (stereo, fusion, utils)."""
    if "ns" in _loaded:
        return _loaded["ns"]
    root = reference_root()
    if root is None:
        raise RuntimeError("CODD reference not found (set CODD_REF or mount /root/reference)")
    for name in ("mmcv", "mmseg"):
        try:
            importlib.import_module(name)
        except ImportError:
            if _SHIM not in sys.path:
                sys.path.insert(0, _SHIM)
    for name in ("pytorch3d", "lietorch", "lietorch_extras"):
        try:
            importlib.import_module(name)
        except ImportError:
            if _SHIM not in sys.path:
                sys.path.insert(0, _SHIM)
    if root not in sys.path:
        sys.path.insert(0, root)

    def _namespace(modname, path):
        m = types.ModuleType(modname)
        m.__path__ = [path]
        sys.modules[modname] = m
        return m

    # `model`, `model.stereo`, `model.motion`, `model.fusion` as bare namespaces: their
    # __init__.py files pull in pytorch3d/lietorch-dependent code at import time.
    _namespace("model", os.path.join(root, "model"))
    for sub in ("stereo", "motion", "fusion", "losses"):
        _namespace(f"model.{sub}", os.path.join(root, "model", sub))
    _namespace("model.stereo.hitnet", os.path.join(root, "model", "stereo", "hitnet"))
    _namespace("model.motion.raft3d", os.path.join(root, "model", "motion", "raft3d"))
    _namespace("model.motion.raft3d.blocks", os.path.join(root, "model", "motion", "raft3d", "blocks"))

    ns = types.SimpleNamespace()
    ns.builder = importlib.import_module("model.builder")
    ns.utils = importlib.import_module("utils")
    ns.warp = importlib.import_module("utils.warp")
    ns.backbone = importlib.import_module("model.stereo.hitnet.backbone")
    ns.initialization = importlib.import_module("model.stereo.hitnet.initialization")
    ns.propagation = importlib.import_module("model.stereo.hitnet.propagation")
    ns.hitnet = importlib.import_module("model.stereo.hitnet.hitnet")
    try:
        ns.fusion = importlib.import_module("model.fusion.fusion")
    except Exception as e:  # pragma: no cover - informative only
        ns.fusion = None
        ns.fusion_error = e
    for attr, mod in (("projective_ops", "model.motion.raft3d.projective_ops"),
                      ("sampler_ops", "model.motion.raft3d.sampler_ops"),
                      ("se3_field", "model.motion.raft3d.se3_field"),
                      ("corr", "model.motion.raft3d.blocks.corr"),
                      ("gru", "model.motion.raft3d.blocks.gru")):
        try:
            setattr(ns, attr, importlib.import_module(mod))
        except Exception as e:  # pragma: no cover - informative only
            setattr(ns, attr, None)
            setattr(ns, attr + "_error", e)
    _loaded["ns"] = ns
    return ns


def build_hitnet(max_disp=192):
    """The reference HITNetMF built through its own registry (configs/models/stereo.py:12-25)."""
    ns = load()
    cfg = dict(
        type="HITNetMF",
        backbone=dict(type="HITUNet"),
        initialization=dict(type="TileInitialization", max_disp=max_disp),
        propagation=dict(type="TilePropagation"),
    )
    model = ns.builder.ESTIMATORS.build(cfg)  # build_estimator adds train_cfg/test_cfg (codd-level only)
    model.eval()
    return model
