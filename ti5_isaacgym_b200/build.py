"""Build libti5step.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m ti5_isaacgym_b200.build [--force]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libti5step.so")
SOURCES = ("ti5_api.cu", "ti5_substep.cu", "ti5_post_physics.cu", "ti5_observe.cu", "ti5_heights.cu", "ti5_gae.cu", "ti5_rollout.cu")
HEADERS = ("ti5_device.cuh", "ti5_host.h", os.path.join("..", "..", "include", "ti5_step.h"))

# -fmad=false: every a*b+c rounds twice, like the separate eager torch ops of the reference.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "csrc", "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log)
    if verbose:
        print(log)
    return LIB


def build_variant(out_path, extra_flags):
    """An experimental / instrumented build next to the product library (e.g. `-DTI5_PROBES` for tools/probe.py);
    load it with TI5_LIB=<out_path>."""
    cmd = [_nvcc(), *[f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")], *extra_flags, "-shared", "-o", out_path] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out_path


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
