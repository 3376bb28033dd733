#!/usr/bin/env python
"""Benchmark of the t1_dh_stand hot path (BASELINE.json metric: env-steps/sec, step math + reward + obs,
8192 envs/GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's algorithm on the host CPU
    torchrun --nproc-per-node N bench.py --gpus N ...               # one rank per GPU, envs sharded 8192/rank

One "step" = one policy step of every env of the rank: 10 x (PD torque + lag push) + post-physics
(derived state, command schedule, termination, 24 reward terms, reset scatter, observation frames +
history rings), plus the GAE reverse scan over the 24-step rollout every 24th step.  Physics is not
part of the metric: the simulator tensors hold one draw of the synthetic near-nominal state
 and `simulate` is a no-op, exactly like the fake gym the reference baseline runs on.

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM and the L2
flushed between timed steps; `e2e` goes through the public `env.step()` with pinned HOST buffers
(H2D of the actions, D2H of reward / reset / time-out flags, a stream sync per step).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 8192
ROLLOUT = 24
GAMMA, LAM = 0.994, 0.9
# algorithmic bytes per env per launch (fp32, D=12, K=47, P=73)
BYTES_SUBSTEP = 584 + 244          # fused torque + lag push of one substep
BYTES_POST = 2412                  # post-physics + reset/observe, ring-view variant
BYTES_ENV_STEP = 10 * BYTES_SUBSTEP + BYTES_POST


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel family, from the committed
    `ncu --set full` capture of this same command (profiles/r01c_ncu_full_summary_8192.json); None if absent."""
    path = os.path.join(ROOT, "profiles", "r01c_ncu_full_summary_8192.json")
    if not os.path.exists(path):
        return None
    want = {"substep": "substep_kernel", "post_physics": "post_physics_kernel", "reset_observe": "reset_observe_kernel",
            "fused_step": "post_physics_kernelILb1", "heights": "heights_kernel"}[kernel]
    table = json.load(open(path))
    rows = next((v for k, v in table.items() if want in k), None)
    if not rows:
        return None
    unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def to_bytes(s):
        v, u = s.split()
        return float(v) * unit[u]
    r = rows[len(rows) // 2]
    return to_bytes(r["dram__bytes_read.sum"]) + to_bytes(r["dram__bytes_write.sum"])


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons through NVML (every 5 ms) while the timed regions run;
    falls back to polling nvidia-smi when pynvml is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.sm_max, self.stop_flag = index, [], set(), None, threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20)):
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        parts = [x.strip() for x in out.strip().split(",")]
        if len(parts) >= 6:
            self.sm.append(float(parts[0]))
            self.sm_max = float(parts[1])
            for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
                if parts[2 + i].lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.005 if self.nvml else 0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unsampled"]}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.sm_max, "samples": len(self.sm),
                "reasons": sorted(self.reasons), "via": "nvml" if self.nvml else "nvidia-smi"}


def make_cfg(num_envs):
    from ti5_isaacgym_b200.envs import DHT1StandCfg
    cfg = DHT1StandCfg()
    cfg.env.num_envs = num_envs
    cfg.terrain.mesh_type = "plane"         # BASELINE config 2: flat plane
    cfg.seed = 5
    return cfg


# ---------------------------------------------------------------------------------------------
# the reference arm / CPU baseline: the oracle port of the reference's torch algorithm on host cores
# ---------------------------------------------------------------------------------------------

def run_cpu_port(num_envs, steps, warmup, seed=1234, device="cpu"):
    from types import SimpleNamespace
    from oracle import t1_oracle as O
    from ti5_isaacgym_b200.envs.t1.t1_robot import robot_constants
    from ti5_isaacgym_b200.sim.synthetic import alloc_sim_tensors, fill_synthetic_state, synthetic_actions
    cfg = make_cfg(num_envs)
    C = O.make_consts(cfg, cfg.sim.dt, robot_constants(cfg), device=device)
    S = O.new_state(C, num_envs)
    gen = torch.Generator().manual_seed(seed)
    sim = alloc_sim_tensors(num_envs, "cpu")
    fill_synthetic_state(sim, S.env_origins.cpu(), gen)
    sim = SimpleNamespace(**{k: v.to(device) for k, v in vars(sim).items()})
    S.episode_length_buf[:] = torch.randint(1, 2000, (num_envs,), generator=gen).to(device)
    S.gait_time[:, 1], S.gait_time[:, 2] = 900, 1500
    actions = synthetic_actions(num_envs, gen, "cpu").to(device)
    pools = {k: v.to(device) for k, v in O.draw_pools(C, num_envs, gen).items()}
    rew = torch.randn(ROLLOUT, num_envs, 1, generator=gen).to(device)
    val = torch.randn(ROLLOUT, num_envs, 1, generator=gen).to(device)
    done = (torch.rand(ROLLOUT, num_envs, 1, generator=gen) < 0.02).byte().to(device)
    last = torch.randn(num_envs, 1, generator=gen).to(device)
    sync = torch.cuda.synchronize if device != "cpu" else (lambda: None)

    def one(i):
        O.step(C, S, sim, actions, pools)
        if (i + 1) % ROLLOUT == 0:
            O.gae_returns(rew, val, done, last, GAMMA, LAM)

    with torch.inference_mode():
        for i in range(warmup):
            one(i)
        sync()
        t0 = time.perf_counter()
        for i in range(steps):
            one(i)
        sync()
        dt = time.perf_counter() - t0
    return num_envs * steps / dt, dt / steps * 1e3


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 for N > 1: the CPU arm runs alone on rank 0 and takes every host core it may use
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, RuntimeError):
        pass
    cores = torch.get_num_threads()
    # bounded sample: the whole run should end within ~2 minutes on the host cores (~30 us per env-step with all
    # threads), so large --steps shrink the number of envs stepped per step (never below 256, at most the workload's 8192)
    budget_env_steps = 3.0e6
    sample_envs = ENVS_PER_GPU
    while sample_envs > 256 and sample_envs * (args.steps + args.warmup) > budget_env_steps:
        sample_envs //= 2
    value, ms = run_cpu_port(sample_envs, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "env-steps/sec (step math + reward + obs)", "value": value,
            "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "t1_dh_stand 8192 envs, flat plane, full step + 24-step rollout GAE",
                                            "device": "host CPU", "physics": "no-op (fake gym)"},
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} steps x {sample_envs} envs of the oracle port (torch CPU, "
                                       "the reference's op chains), all host threads"},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------

def cuda_arm(args):
    import torch.distributed as dist
    from ti5_isaacgym_b200.algo.rollout_storage import gae_returns_, make_gae_scratch
    from ti5_isaacgym_b200.envs import T1DHStandEnv
    from ti5_isaacgym_b200.sim.synthetic import SimParams, fill_synthetic_state, synthetic_actions

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the step math has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
        group = dist.group.WORLD
    N = args.envs
    cfg = make_cfg(N)
    cfg.seed = 5 + rank
    env = T1DHStandEnv(cfg, SimParams(dt=cfg.sim.dt), 1, dev, True, rng_mode="philox", div_mode="reciprocal",
                       use_cuda_graph=True, materialize_obs=args.materialize)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    fill_synthetic_state(env.gym.tensors, env.env_origins, gen)
    env.reset()
    env.episode_length_buf = torch.randint(1, 2000, (N,), generator=gen, device=dev)
    actions = synthetic_actions(N, gen, dev)
    T = ROLLOUT
    rew = torch.randn(T, N, 1, generator=gen, device=dev)
    val = torch.randn(T, N, 1, generator=gen, device=dev)
    done = (torch.rand(T, N, 1, generator=gen, device=dev) < 0.02).byte()
    last = torch.randn(N, 1, generator=gen, device=dev)
    ret, adv = torch.empty_like(rew), torch.empty_like(rew)
    scratch = make_gae_scratch(N, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one(i):
        env.step(actions)
        if (i + 1) % T == 0:
            gae_returns_(rew, val, done, last, ret, adv, GAMMA, LAM, scratch, group)

    for i in range(max(args.warmup, 3)):
        one(i)
    # ---- device-timed region: K steps, L2 flushed (outside the event brackets) between steps ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        starts[i].record()
        one(i)
        ends[i].record()
    barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = world * N * args.steps / (dev_ms * 1e-3)

    # ---- per-launch duration of the dominant kernel family (events on the launching stream) -------
    kt = kernel_times(env, actions, steps=min(args.steps, 48))

    # ---- end to end through env.step() with pinned host buffers ------------------------------------
    # pinned, device-mapped host buffers wired into the step's CUDA graph (LeggedRobot.enable_host_io): every
    # step_host() has the first substep kernel read the actions from host memory, runs the step, has the observation
    # kernel store [rew | reset | time_outs] into host memory, and waits for the stream
    h_act, h_out = env.enable_host_io()
    h_act.copy_(actions.cpu())
    for i in range(3):
        env.step_host()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        env.step_host()
        if (i + 1) % T == 0:
            gae_returns_(rew, val, done, last, ret, adv, GAMMA, LAM, scratch, group)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * N * args.steps / float(t.item())
    clocks = sampler.summary() if sampler else None      # sampled across the device-timed, per-phase and e2e regions

    if rank == 0:
        peak, peak_src = peaks()
        dom = max(kt, key=lambda k: kt[k]["share_ms"])
        # algorithmic bytes of one launch of the dominant family (SURVEY 8d figure x envs per launch);
        # the post phase's 2412 B/env are spread over its two launches
        bytes_per_launch = kt[dom]["bytes_per_launch"]
        ach = bytes_per_launch / (kt[dom]["ms_per_launch"] * 1e-3) / 1e9
        launches_per_step = env.launches_per_step
        cpu = torch_gpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, ms = run_cpu_port(4096, 40, 3)
            cpu = {"value": v, "unit": "env-steps/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": "40 steps x 4096 envs (BASELINE config 1) of the oracle port, torch CPU, all host threads, "
                             f"{ms:.1f} ms/step"}
            # the reference's eager-torch algorithm on THIS GPU (the north_star's 20x denominator); context only
            v, ms = run_cpu_port(N, 24, 3, device=dev)
            torch_gpu = {"value": v, "unit": "env-steps/s", "ms_per_step": ms,
                         "sample": f"24 steps x {N} envs of the oracle port in eager torch on the same B200"}
        rollout = rollout_bench(env, gen, actions) if world == 1 and not args.no_rollout and not args.materialize else None
        sweep = None
        if world == 1 and not args.no_sweep and N == ENVS_PER_GPU and not args.materialize:
            del flush
            sweep = [sweep_point(n, dev) for n in (1024, 16384, 65536)]
        line = {
            "metric": "env-steps/sec (step math + reward + obs)", "value": value, "unit": "env-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"t1_dh_stand {N} envs/GPU, flat plane, full fused step + 24-step rollout GAE",
                       "envs_per_gpu": N, "frame_stack": cfg.env.frame_stack, "obs": "materialised" if args.materialize else "ring view",
                       "rng": "in-kernel Philox4x32-10", "physics": "no-op (synthetic state, SURVEY 8d)",
                       "l2": "flushed between timed steps (256 MiB write outside the event brackets)",
                       "launch": f"one CUDA graph per step ({env.launches_per_step} kernels"
                                 + (", programmatic dependent launches)" if env._chain_launches else ")")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h_act.numel() * 4 * world,
                    "d2h_bytes_per_step": h_out.numel() * world,
                    "note": "env.step_host(): actions read from pinned host memory by the first kernel of the step, reward/reset/time-out flags stored into pinned host memory by the last one (mapped pages, no copy-engine hop), stream sync every step; L2 not flushed"},
            "gpu_launches": launches_per_step * args.steps + 2 * (args.steps // T),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": ncu_traffic(dom) if N == ENVS_PER_GPU else None, "peak_source": peak_src, "bytes_per_launch": bytes_per_launch,
                         "ms_per_launch": kt[dom]["ms_per_launch"],
                         "whole_step": {"bytes": BYTES_ENV_STEP * N, "achieved": BYTES_ENV_STEP * N * args.steps / (dev_ms * 1e-3) / 1e9},
                         "kernels": kt,
                         "how": "CUDA events on the launching stream around one CUDA graph per kernel family (L2 flushed before each step); per-launch = family time / launches in the family; algorithmic bytes = SURVEY 8d per-env figure x envs (post phase split over its two kernels, see bench.py)"},
            "cpu_baseline": cpu,
            "reference_torch_gpu": torch_gpu,
            "rollout_storage": rollout,
            "sweep": sweep,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sweep_point(N, dev, steps=48, warmup=12):
    """One point of BASELINE config 5 (num_envs sweep): the same device-timed step (L2 flushed before every step, GAE
    every 24th) and per-phase graph timing as the main line, on a fresh env of N envs."""
    from ti5_isaacgym_b200.algo.rollout_storage import gae_returns_, make_gae_scratch
    from ti5_isaacgym_b200.envs import T1DHStandEnv
    from ti5_isaacgym_b200.sim.synthetic import SimParams, fill_synthetic_state, synthetic_actions
    cfg = make_cfg(N)
    env = T1DHStandEnv(cfg, SimParams(dt=cfg.sim.dt), 1, dev, True, rng_mode="philox", div_mode="reciprocal", use_cuda_graph=True,
                       materialize_obs=False)
    gen = torch.Generator(device=dev).manual_seed(4321)
    fill_synthetic_state(env.gym.tensors, env.env_origins, gen)
    env.reset()
    env.episode_length_buf = torch.randint(1, 2000, (N,), generator=gen, device=dev)
    actions = synthetic_actions(N, gen, dev)
    T = ROLLOUT
    rew = torch.randn(T, N, 1, generator=gen, device=dev)
    val = torch.randn(T, N, 1, generator=gen, device=dev)
    done = (torch.rand(T, N, 1, generator=gen, device=dev) < 0.02).byte()
    last = torch.randn(N, 1, generator=gen, device=dev)
    ret, adv = torch.empty_like(rew), torch.empty_like(rew)
    scratch = make_gae_scratch(N, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def one(i):
        env.step(actions)
        if (i + 1) % T == 0:
            gae_returns_(rew, val, done, last, ret, adv, GAMMA, LAM, scratch, None)
    for i in range(warmup):
        one(i)
    marks = []
    for i in range(steps):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); one(i); b.record()
        marks.append((a, b))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in marks) / steps
    kt = kernel_times(env, actions, steps=24)
    peak, _ = peaks()
    out = {"envs_per_gpu": N, "value": N / (ms * 1e-3), "unit": "env-steps/s", "ms_per_step": ms,
           "kernels": {k: {"ms_per_launch": v["ms_per_launch"], "achieved": v["achieved_GBps"], "frac": v["achieved_GBps"] / peak}
                       for k, v in kt.items()},
           "whole_step": {"bytes": BYTES_ENV_STEP * N, "achieved": BYTES_ENV_STEP * N / (ms * 1e-3) / 1e9,
                          "frac": BYTES_ENV_STEP * N / (ms * 1e-3) / 1e9 / peak}}
    del env, flush
    torch.cuda.empty_cache()
    return out


def rollout_bench(env, gen, actions):
    """SURVEY 8(f) rows 1-2 on the same env: per-step cost of recording a transition and per-mini-batch cost of
    drawing one, frame-log storage (ti5_store_transition / ti5_gather_minibatch) next to the reference's storage
    algorithm (rs:59-74 copies, rs:152-164 index gathers) as eager torch on the same GPU.  Device-timed with CUDA
    events; L2 flushed before every timed gather."""
    from ti5_isaacgym_b200.algo.rollout_storage import FrameLogRolloutStorage, RolloutStorage
    N, dev, T, NMB = env.num_envs, env.device, ROLLOUT, 4
    ev = lambda: torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fl = FrameLogRolloutStorage(env, T)
    plain = RolloutStorage(N, T, [env.num_obs], [env.num_privileged_obs], [12], None, dev)
    obs, priv = env.get_observations(), env.get_privileged_observations()
    mk = lambda: RolloutStorage.Transition()
    val, logp, sig = torch.randn(N, 1, device=dev), torch.randn(N, device=dev), torch.full((N, 12), 0.3, device=dev)
    t_env, t_fl, t_pl = [], [], []
    for rollout in range(3):                     # first pass = warm-up (graph re-capture with the log, allocator)
        fl.clear(); plain.clear()
        for t in range(T):
            tr = mk()
            tr.actions, tr.values, tr.actions_log_prob, tr.action_mean, tr.action_sigma = actions, val, logp, actions, sig
            tr.observations, tr.critic_observations = obs, priv
            held_obs, held_priv = obs.clone(), priv.clone()          # what the reference's storage copies one step late
            e = [ev() for _ in range(4)]
            e[0].record()
            obs, priv, rew, dones, infos = env.step(actions)
            e[1].record()
            fl.store_step(tr, rew, dones, infos["time_outs"], GAMMA)
            e[2].record()
            tp = mk()                                                  # dh_ppo.py:93-103 + rs:59-74 as the reference runs them
            tp.actions, tp.values, tp.actions_log_prob, tp.action_mean, tp.action_sigma = actions, val, logp, actions, sig
            tp.observations, tp.critic_observations = held_obs, held_priv
            tp.rewards = rew.clone()
            tp.dones = dones
            tp.rewards += GAMMA * torch.squeeze(tp.values * infos["time_outs"].unsqueeze(1), 1)
            plain.add_transitions(tp)
            e[3].record()
            if rollout:
                t_env.append((e[0], e[1])); t_fl.append((e[1], e[2])); t_pl.append((e[2], e[3]))
    torch.cuda.synchronize()
    mean_ms = lambda pairs: statistics.mean(a.elapsed_time(b) for a, b in pairs)
    mb = T * N // NMB
    idx = torch.randperm(NMB * mb, device=dev, generator=gen)
    cols = [t.flatten(0, 1) for t in (plain.observations, plain.privileged_observations, plain.actions, plain.values,
                                      plain.advantages, plain.returns, plain.actions_log_prob, plain.mu, plain.sigma)]
    g_fl, g_pl = [], []
    for rep in range(3):
        for i in range(NMB):
            sl = idx[i * mb:(i + 1) * mb]
            flush.fill_(i)
            a, b = ev(), ev()
            a.record(); out = fl.gather(sl); b.record()
            flush.fill_(i + 1)
            c, d = ev(), ev()
            c.record(); ref = [col[sl] for col in cols]; d.record()
            if rep:
                g_fl.append((a, b)); g_pl.append((c, d))
            if rep == 0 and i == 0:
                same = all(torch.equal(out[k], r) for k, r in zip(("obs", "critic_obs", "actions", "values"), ref))
            del out, ref
    torch.cuda.synchronize()
    row_bytes = 4 * (env.num_obs + env.num_privileged_obs + 3 * 12 + 4)
    gather_ms = mean_ms(g_fl)
    logs = env.frame_logs()
    kept = sum(t.numel() * t.element_size() for t in (logs.frame_log, logs.priv_log, logs.valid_log))
    return {
        "rollout": f"T={T} steps x {N} envs, {NMB} mini-batches of {mb} rows (t1_cfg:455-468)",
        "env_step_with_frame_log_ms": mean_ms(t_env),
        "store": {"ours_us_per_step": 1e3 * mean_ms(t_fl), "launches_per_step": 1,
                  "reference_algorithm_torch_gpu_us_per_step": 1e3 * mean_ms(t_pl),
                  "reference_bytes_per_step": 2 * 4 * N * (env.num_obs + env.num_privileged_obs),
                  "ours_bytes_per_step": N * (4 * (12 * 6 + 8) + 2)},
        "gather": {"ours_ms_per_minibatch": gather_ms, "reference_algorithm_torch_gpu_ms_per_minibatch": mean_ms(g_pl),
                   "bytes_written_per_minibatch": mb * row_bytes,
                   "achieved_GBps_read_plus_write": 2 * mb * row_bytes / (gather_ms * 1e-3) / 1e9,
                   "matches_reference_algorithm": bool(same), "l2": "flushed before every timed gather"},
        "observation_storage_bytes": {"ours_frame_logs": kept,
                                      "reference": 4 * T * N * (env.num_obs + env.num_privileged_obs)},
    }


# SURVEY 8(d) row B (2412 B per env-step for the whole post phase) split over its two kernels: ti5_post_physics reads
# every input of the phase once (922 B) and writes the derived state, counters, reward and episode sums (266 B);
# ti5_reset_observe writes both frames into both mirrored ring slots (960 B), the last_* copies and the reference
# pose (264 B) — its re-reads of what ti5_post_physics produced are not counted.
BYTES_POST_PHYSICS = 922 + 266
BYTES_RESET_OBSERVE = BYTES_POST - BYTES_POST_PHYSICS
# ti5_fused_step (clip + 10 substeps + post-physics in one launch): what has to move once the substeps share one pass
# over an env — reads: actions 48, five actuator arrays 240, lag index + stamp 12, on average 8.2
# lagged action rows that predate the step 394, the post-physics inputs 922 (joint state included); writes: ten action /
# DOF / IMU ring rows 1680, actions + final torques + multipliers 144, the post-physics outputs 266.
BYTES_FUSED_STEP = (48 + 240 + 12 + 394 + 922) + (1680 + 144 + 266)
BYTES_HEIGHTS = 1122 + 748          # 3 x 187 int16 gathers + (N,187) fp32 out
KERNEL_BYTES = {"substep": BYTES_SUBSTEP, "post_physics": BYTES_POST_PHYSICS, "reset_observe": BYTES_RESET_OBSERVE,
                "fused_step": BYTES_FUSED_STEP, "heights": BYTES_HEIGHTS}


def kernel_times(env, actions, steps):
    """Average device time of the kernel families of a step, CUDA events on the launching stream around the replay of
    one CUDA graph per family (a single whole-step graph cannot be bracketed inside; the event records also cut the
    programmatic launch chain at the family boundaries, so the times add up to a little more than the whole-step
    time).  Per-launch time = family time / launches in the family."""
    graphs = env.capture_phase_graphs()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=env.device)
    env._actions_in.copy_(actions)
    acc = {name: [] for name, _, _ in graphs}
    marks = []
    for i in range(steps):
        flush.fill_(i & 0xFF)
        e = [ev() for _ in range(len(graphs) + 1)]
        e[0].record()
        for k, (_, g, _) in enumerate(graphs):
            g.replay()
            e[k + 1].record()
        env._finish_step()
        marks.append(e)
    torch.cuda.synchronize()
    for e in marks:
        for k, (name, _, _) in enumerate(graphs):
            acc[name].append(e[k].elapsed_time(e[k + 1]))
    out = {}
    for name, _, n in graphs:
        phase = statistics.mean(acc[name])
        extra = env._params.priv_frame - 73 if name == "reset_observe" else 0       # measured heights ride in the critic frame
        nbytes = (KERNEL_BYTES[name] + 4 * 3 * extra) * env.num_envs
        out[name] = {"phase_ms": phase, "launches": n, "ms_per_launch": phase / n, "share_ms": phase,
                     "bytes_per_launch": nbytes, "achieved_GBps": nbytes / (phase / n * 1e-3) / 1e9}
    del flush
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=240)
    ap.add_argument("--warmup", type=int, default=24)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU (BASELINE metric: 8192)")
    ap.add_argument("--materialize", action="store_true", help="also write contiguous (N,3102)/(N,219) observations")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rollout", action="store_true", help="skip the rollout-storage measurement (SURVEY 8f rows 1-2)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the num_envs sweep points (BASELINE config 5)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        cuda_arm(args)


if __name__ == "__main__":
    main()
