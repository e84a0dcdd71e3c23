"""Nested-class configuration objects (mirror of reference legged_gym/envs/base/base_config.py:33-55).

Behaviour kept: instantiating a config instantiates every nested class found on it, recursively,
and stores the instance on the *instance* under the same name, so ``cfg.env.num_envs`` reads and
writes per-instance state while ``Cfg.env`` stays a class users can subclass.
"""
import inspect


class BaseConfig:
    def __init__(self) -> None:
        self.init_member_classes(self)

    @staticmethod
    def init_member_classes(obj):
        for name in dir(obj):
            if name == "__class__":
                continue
            member = getattr(obj, name)
            if inspect.isclass(member):
                inst = member()
                setattr(obj, name, inst)
                BaseConfig.init_member_classes(inst)


def cfg_from_spec(name, bases, spec, module=None):
    """Build a config class from a nested dict: dict values become nested classes (inheriting from
    the same-named nested class of the first base that has one), everything else a class attribute.
    ``{"__replace__": True, ...}`` makes a nested class that does NOT inherit."""
    ns = {}
    for key, val in spec.items():
        if isinstance(val, dict) and val.get("__cfgclass__", True) and not val.get("__value__", False):
            val = dict(val)
            replace = val.pop("__replace__", False)
            val.pop("__cfgclass__", None)
            parent = ()
            if not replace:
                for b in bases:
                    p = getattr(b, key, None)
                    if inspect.isclass(p):
                        parent = (p,)
                        break
            ns[key] = cfg_from_spec(key, parent, val, module)
        elif isinstance(val, dict):
            val = dict(val)
            val.pop("__value__", None)
            ns[key] = val
        else:
            ns[key] = val
    if module is not None:
        ns["__module__"] = module
    return type(name, tuple(bases), ns)


def value(d):
    """Mark a dict as a plain value (e.g. default_joint_angles) rather than a nested class."""
    d = dict(d)
    d["__value__"] = True
    return d
