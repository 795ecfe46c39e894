import inspect
import os


class Registry:
    """name -> class table with mmcv's ``register_module`` / ``build`` calls."""

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

        if module is not None:
            return _register(module)
        return _register

    def build(self, cfg, default_args=None):
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        obj_type = args.pop("type")
        cls = obj_type if inspect.isclass(obj_type) else self.module_dict.get(obj_type)
        if cls is None:
            raise KeyError(f"{obj_type} is not in the {self.name} registry")
        return cls(**args)


def mkdir_or_exist(dir_name, mode=0o777):
    if dir_name:
        os.makedirs(os.path.expanduser(dir_name), mode=mode, exist_ok=True)
