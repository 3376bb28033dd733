def __getattr__(name):
    raise RuntimeError("matplotlib is stubbed in the oracle shim")
