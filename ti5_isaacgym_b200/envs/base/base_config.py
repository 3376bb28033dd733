"""Config plumbing with the reference's attribute API.

The reference describes its configuration as nested Python classes that are
instantiated recursively (`humanoid/envs/base/base_config.py:9-25`).  Callers
(`task_registry`, the PPO runner, the env) only ever do attribute access and
`class_to_dict`, so here the class trees are *generated* from plain dict
specifications (`build_cfg`) instead of being spelled out as class bodies.
"""
import inspect


class BaseConfig:
    """Instantiating a config turns every nested class attribute into an instance."""

    def __init__(self):
        self.init_member_classes(self)

    @staticmethod
    def init_member_classes(obj):
        """Public in the reference (base_config.py:14); kept so the config dumps agree."""
        _instantiate_members(obj)


def _instantiate_members(node):
    for name in dir(node):
        if name == "__class__":
            continue
        member = getattr(node, name)
        if inspect.isclass(member):
            inst = member()
            setattr(node, name, inst)
            _instantiate_members(inst)


class ns(dict):
    """Marks a dict in a spec as a *nested config class* (plain dicts stay data)."""


class ns_new(ns):
    """A nested config class that does NOT inherit from the base's same-named class
    (the reference's bare `class rewards:` / `class normalization:` in t1_cfg:360,416)."""


def build_cfg(name, spec, base=None, root_base=BaseConfig):
    """Create class `name` from `spec`; nested `ns` entries become nested classes that
    inherit from the same-named nested class of `base` (the reference's
    `class env(LeggedRobotCfg.env)` idiom, t1_dh_stand_config.py:9)."""
    body = {}
    for key, val in spec.items():
        if isinstance(val, ns):
            inherit = base is not None and not isinstance(val, ns_new)
            parent = getattr(base, key, None) if inherit else None
            body[key] = build_cfg(key, val, parent, root_base=None)
        else:
            body[key] = val
    if base is not None:
        bases = (base,)
    elif root_base is not None:
        bases = (root_base,)
    else:
        bases = ()
    return type(name, bases, body)
