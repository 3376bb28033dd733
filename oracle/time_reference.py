"""Time the UNMODIFIED reference (`/root/reference`, under `oracle/shim`) next to the oracle port on the host CPU
of the build container, same env count, same synthetic simulator tensors, own RNG (torch's generator, no pools).
TEST INFRASTRUCTURE — only usable where the reference tree is mounted; the GPU box times the port alone
(`bench.py --impl reference`).  Answers round-1 verdict weak #10: how does the port, which stands in for the
reference in the bench's reference arm, compare with the reference itself?

    python oracle/time_reference.py [--envs 8192] [--steps 10] [--warmup 2] [--config 2|3] [--out profiles/...json]
"""
import argparse
import json
import os
import sys
import time
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import t1_oracle as O                                   # noqa: E402
from oracle.reference_driver import ReferenceDriver, adopt_reference_state   # noqa: E402
from oracle.pin_against_reference import robot_from_env                     # noqa: E402
from ti5_isaacgym_b200.sim.synthetic import fill_synthetic_state            # noqa: E402


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--out", default=None)
    ap.add_argument("--config", type=int, default=2, choices=[2, 3],
                    help="BASELINE config: 2 = plane; 3 = trimesh + measured heights + push_robots + 5 %% terminations per step")
    a = ap.parse_args()
    N = a.envs
    terrain = heights = None
    if a.config == 3:
        def edit(c):
            c.terrain.measure_heights = True
            c.env.num_privileged_obs = 3 * (73 + 187)
            c.domain_rand.push_robots = True
        drv = ReferenceDriver(N, mesh_type="trimesh", cfg_edit=edit, seed=11)
    else:
        drv = ReferenceDriver(N, mesh_type="plane", seed=11)
    env = drv.env
    if a.config == 3:
        terrain = SimpleNamespace(env_length=env.terrain.env_length, max_level=env.max_terrain_level, origins=env.terrain_origins)
        heights = env.height_samples
    contact_rate = 0.05 if a.config == 3 else 0.01
    gen = torch.Generator().manual_seed(1234)
    fill_synthetic_state(drv.sim, env.env_origins, gen, base_contact_rate=contact_rate)
    env.reset()
    env.episode_length_buf[:] = torch.randint(1, 2000, (N,), generator=gen)
    env.phase_length_buf[:] = env.episode_length_buf
    C = O.make_consts(drv.cfg, drv.cfg.sim.dt, robot_from_env(env), terrain=terrain)
    S = adopt_reference_state(O.new_state(C, N), env)
    sim = drv.sim
    actions = [torch.randn(N, 12, generator=gen) for _ in range(a.warmup + a.steps)]

    def timed(fn):
        for i in range(a.warmup):
            fn(actions[i])
        t0 = time.perf_counter()
        for i in range(a.steps):
            fn(actions[a.warmup + i])
        return (time.perf_counter() - t0) / a.steps

    t_ref = timed(lambda act: env.step(act.clone()))

    def port_step(act):
        pools = O.draw_pools(C, N, gen)            # the port takes its uniforms as inputs: drawing them is part of its step
        O.step(C, S, sim, act, pools, terrain=terrain, height_samples=heights)

    t_port = timed(port_step)
    line = {
        "what": "unmodified reference (T1DHStandEnv.step under oracle/shim, fake gym: simulate is a no-op) vs the oracle port, "
                "host CPU of the build container",
        "baseline_config": a.config, "envs": N, "steps": a.steps, "warmup": a.warmup, "threads": torch.get_num_threads(), "cpu_model": cpu_model(),
        "reference_ms_per_step": t_ref * 1e3, "reference_env_steps_per_s": N / t_ref,
        "port_ms_per_step": t_port * 1e3, "port_env_steps_per_s": N / t_port,
        "port_over_reference": t_ref / t_port,
    }
    print(json.dumps(line))
    if a.out:
        with open(a.out, "w") as f:
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
