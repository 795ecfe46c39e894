"""Registry used by the drop-in modules.

The reference registers its networks in mmseg's ``MODELS`` registry and builds them from config
dicts by ``type=`` string (model/builder.py:5-21, model/codd.py:44-54).  When mmseg is installed
the codd_b200 modules register in *that* registry (``force=True``: same names, so a reference
config builds the B200 implementation).  When mmseg / mmcv are absent — as in this image — a
minimal registry with the same ``register_module`` / ``build`` calls is used instead.
"""
import inspect

try:  # pragma: no cover - not installed in this image
    from mmseg.models.builder import BACKBONES, MODELS  # type: ignore
    HAVE_MMSEG = True
except Exception:  # noqa: BLE001
    HAVE_MMSEG = False

    class Registry:
        def __init__(self, name):
            self.name = name
            self.module_dict = {}

        def get(self, key):
            return self.module_dict.get(key)

        def register_module(self, name=None, force=False, module=None):
            def _register(cls):
                key = name or cls.__name__
                if key in self.module_dict and not force:
                    raise KeyError(f"{key} is already registered in {self.name}")
                self.module_dict[key] = cls
                return cls

            return _register(module) if module is not None else _register

        def build(self, cfg, default_args=None):
            if not isinstance(cfg, dict) or "type" not in cfg:
                raise TypeError("cfg must be a dict with a 'type' key")
            args = dict(cfg)
            for k, v in (default_args or {}).items():
                args.setdefault(k, v)
            t = args.pop("type")
            cls = t if inspect.isclass(t) else self.module_dict.get(t)
            if cls is None:
                raise KeyError(f"{t} is not in the {self.name} registry")
            return cls(**args)

    MODELS = Registry("models")
    BACKBONES = MODELS

ESTIMATORS = MODELS  # model/builder.py:7
LOSSES = MODELS


def register(registry=MODELS):
    """``@register()``: mmseg-compatible registration that overrides a reference class of the
    same name when both are importable."""
    return registry.register_module(force=True)


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)
