"""gymutil subset used by base_task.py:21 and helpers.py."""


def parse_device_str(device_str):
    s = str(device_str).lower()
    if s in ("cpu", "cuda"):
        return s, 0
    kind, idx = s.split(":")
    return kind, int(idx)


def parse_sim_config(cfg, sim_params):
    for k, v in cfg.items():
        if isinstance(v, dict):
            sub = getattr(sim_params, k, None)
            if sub is not None:
                for kk, vv in v.items():
                    setattr(sub, kk, vv)
        else:
            setattr(sim_params, k, v)


def parse_arguments(description="", custom_parameters=()):
    raise RuntimeError("CLI parsing is not part of the shim")
