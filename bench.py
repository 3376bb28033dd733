#!/usr/bin/env python
"""Benchmark of the t1_dh_stand hot path (BASELINE.json metric: env-steps/sec, step math + reward + obs,
8192 envs/GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path, BASELINE config 2
    python bench.py --config 3 ...                                  # trimesh + measured heights + pushes + 5 % resets
    python bench.py --impl reference [--config 2|3] ...             # the reference's algorithm on the host CPU
    python bench.py --ppo-iter [--gpus N]                           # BASELINE config 4: a whole PPO iteration with the
                                                                    # gradient all-reduce (one JSON line of its own)
    torchrun --nproc-per-node N bench.py --gpus N ...               # one rank per GPU, envs sharded 8192/rank

One "step" = one policy step of every env of the rank: action clip + 10 x (PD torque + lag pushes) + post-physics
(derived state, command schedule, termination, 24 reward terms, reset scatter, observation frames + history rings),
plus the GAE reverse scan over the 24-step rollout on every 24th step (the step counter runs through warm-up, so the
cadence does not depend on --steps).  Physics is not part of the metric: the simulator tensors hold one draw of the
synthetic near-nominal state and `simulate` is a no-op, exactly like the fake gym the reference baseline
runs on.

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM and the L2 flushed between timed
steps; `e2e` goes through the public `env.step_host()` with pinned HOST buffers (actions in, reward / reset / time-out
flags out, a stream sync per step).  The benchmarked mode (in-kernel Philox, CUDA graph, chained launches, fused step)
is checked in tests/ against itself (graph == direct launches, fused == 12 launches, chained == plain, bit-exact) and
statistically; parity with the reference is established in pools mode on the same kernels.
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
# ---- algorithmic bytes per env (fp32, D=12, K=47, P=73, DEC=10), SURVEY.md 8(d) / DESIGN.md section 4 ------------
BYTES_SUBSTEP = 584 + 244          # torque + lag push of one substep as separate launches (12-launch sequence)
BYTES_POST = 2412                  # post-physics + reset/observe, ring-view variant
BYTES_ENV_STEP = 10 * BYTES_SUBSTEP + BYTES_POST        # 10 692 B per env-step
# SURVEY 8(d) row B (2412 B per env-step for the whole post phase) split over its two kernels: ti5_post_physics reads
# every input of the phase once (922 B) and writes the derived state, counters, reward and episode sums (266 B);
# ti5_reset_observe writes both frames into both mirrored ring slots (960 B), the last_* copies and the reference
# pose (264 B) — its re-reads of what ti5_post_physics produced are not counted.
BYTES_POST_PHYSICS = 922 + 266
BYTES_RESET_OBSERVE = BYTES_POST - BYTES_POST_PHYSICS
# ti5_fused_step (clip + 10 substeps + post-physics in ONE launch): what has to move once the substeps share one pass
# over an env — reads: actions 48, five actuator arrays 240, lag index + stamp 12, on average 8.2 lagged action rows
# that predate the step 394, the post-physics inputs 922 (joint state included); writes: ten action / DOF / IMU ring
# rows 1680, the ten substeps' torques 480, actions + final torques + multipliers 144, the post-physics outputs 266.
# (SURVEY's per-launch figures would credit this kernel with 10 x 828 + 1188 = 9468 B; the smaller number is used.)
BYTES_FUSED_STEP = (48 + 240 + 12 + 394 + 922) + (1680 + 480 + 144 + 266)
BYTES_HEIGHTS = 1122 + 748         # 3 x 187 int16 gathers + (N,187) fp32 out
BYTES_ENV_STEP_FUSED = BYTES_FUSED_STEP + BYTES_RESET_OBSERVE
KERNEL_BYTES = {"substep": BYTES_SUBSTEP, "post_physics": BYTES_POST_PHYSICS, "reset_observe": BYTES_RESET_OBSERVE,
                "fused_step": BYTES_FUSED_STEP, "heights": BYTES_HEIGHTS}
NCU_SUMMARY = os.path.join(ROOT, "profiles", "r02_ncu_full_summary_8192.json")


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel family, from the committed
    `ncu --set full` capture of this same step (profiles/r02_ncu_full_summary_8192.json, written by
    tools/ncu_summary.py); None if absent."""
    if not os.path.exists(NCU_SUMMARY):
        return None
    want = {"substep": "substep_kernel", "post_physics": "post_physics_kernel", "reset_observe": "reset_observe_kernel",
            "fused_step": "post_physics_kernel", "heights": "heights_kernel"}[kernel]
    table = json.load(open(NCU_SUMMARY))
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


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons through NVML (every 5 ms) while the timed regions run;
    falls back to polling nvidia-smi when pynvml is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.sm_max, self.stop_flag = index, [], set(), None, threading.Event()
        self.interval = 0.005          # seconds between NVML samples (the wall-clock e2e region widens it, see cuda_arm)
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
            self.stop_flag.wait(self.interval if self.nvml else 0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unsampled"]}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.sm_max, "samples": len(self.sm),
                "reasons": sorted(self.reasons), "via": "nvml" if self.nvml else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------
# workloads: BASELINE.json configs 2 (plane) and 3 (trimesh + measured heights + pushes + frequent resets)
# ---------------------------------------------------------------------------------------------
CONFIG_NOTE = {2: "flat plane", 3: "trimesh terrain, measured heights (187 points, 260-float privileged frame), "
                                   "push_robots, 5 % base-contact terminations per step"}
CONTACT_RATE = {2: 0.01, 3: 0.05}
COUNTER0 = {2: 0, 3: 240000}        # config 3: common_step_counter where the push window is 20 of 600 steps (t1_cfg:190-192)


def make_cfg(num_envs, config=2, frame_stack=66):
    from ti5_isaacgym_b200.envs import make_t1_cfg
    cfg = make_t1_cfg(frame_stack=frame_stack)()
    cfg.env.num_envs = num_envs
    cfg.terrain.mesh_type = "plane"
    if config == 3:
        cfg.terrain.mesh_type = "trimesh"
        cfg.terrain.measure_heights = True
        cfg.env.num_privileged_obs = cfg.env.c_frame_stack * (cfg.env.single_num_privileged_obs + cfg.terrain.num_height)
        cfg.domain_rand.push_robots = True
    cfg.seed = 5
    return cfg


def workload_name(N, config, H):
    return (f"t1_dh_stand {N} envs/GPU, {CONFIG_NOTE[config]}, frame_stack {H}, full step + 24-step rollout GAE "
            "(every 24th step)")


class StepLoop:
    """The env, its inputs, and the GAE tensors of one rank; `one()` is one step of the metric."""

    def __init__(self, N, dev, config=2, frame_stack=66, rank=0, graph=True, materialize=False, group=None, seed=1234):
        from ti5_isaacgym_b200.algo.rollout_storage import make_gae_scratch
        from ti5_isaacgym_b200.envs import T1DHStandEnv
        from ti5_isaacgym_b200.sim.synthetic import SimParams, fill_synthetic_state, synthetic_actions
        self.cfg = cfg = make_cfg(N, config, frame_stack)
        cfg.seed = 5 + rank
        self.N, self.dev, self.group, self.config = N, dev, group, config
        self.env = env = T1DHStandEnv(cfg, SimParams(dt=cfg.sim.dt), 1, dev, True, rng_mode="philox", div_mode="reciprocal",
                                      use_cuda_graph=graph, materialize_obs=materialize)
        gen = self.gen = torch.Generator(device=dev).manual_seed(seed + rank)
        fill_synthetic_state(env.gym.tensors, env.env_origins, gen, base_contact_rate=CONTACT_RATE[config])
        env.reset()
        env.episode_length_buf = torch.randint(1, 2000, (N,), generator=gen, device=dev)
        if COUNTER0[config]:
            env.set_common_step_counter(COUNTER0[config])
        self.actions = synthetic_actions(N, gen, dev)
        T = ROLLOUT
        self.rew = torch.randn(T, N, 1, generator=gen, device=dev)
        self.val = torch.randn(T, N, 1, generator=gen, device=dev)
        self.done = (torch.rand(T, N, 1, generator=gen, device=dev) < 0.02).byte()
        self.last = torch.randn(N, 1, generator=gen, device=dev)
        self.ret, self.adv = torch.empty_like(self.rew), torch.empty_like(self.rew)
        self.scratch = make_gae_scratch(N, dev)
        self.count = 0          # env steps taken so far: the GAE cadence runs through warm-up and every timed region
        self.gae_calls = 0

    def gae(self):
        from ti5_isaacgym_b200.algo.rollout_storage import gae_returns_
        gae_returns_(self.rew, self.val, self.done, self.last, self.ret, self.adv, GAMMA, LAM, self.scratch, self.group)
        self.gae_calls += 1

    def warm_gae(self):
        """One untimed GAE call during warm-up: with `--steps 20 --warmup 5` the cadence puts the FIRST call of the process
        inside the timed region, and a first call pays the lazy load of its two kernels (150 us: 46.2 instead of 38.5 us
        per step over 20 steps).  The cadence itself is not touched."""
        self.gae()
        self.gae_calls -= 1

    def one(self, host=False):
        self.env.step_host() if host else self.env.step(self.actions)
        self.count += 1
        if self.count % ROLLOUT == 0:
            self.gae()

    def timed(self, steps, flush):
        """K steps, device-timed: CUDA events on the launching stream around every step, the L2 flushed (256 MiB write)
        outside the brackets before each.  Returns (total ms, GAE calls inside)."""
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        g0 = self.gae_calls
        for i in range(steps):
            flush.fill_(i & 0xFF)
            starts[i].record()
            self.one()
            ends[i].record()
        torch.cuda.synchronize()
        return sum(s.elapsed_time(e) for s, e in zip(starts, ends)), self.gae_calls - g0


# ---------------------------------------------------------------------------------------------
# the reference arm / CPU baseline: the oracle port of the reference's torch algorithm on host cores
# ---------------------------------------------------------------------------------------------

def run_cpu_port(num_envs, steps, warmup, seed=1234, device="cpu", config=2):
    from types import SimpleNamespace
    from oracle import t1_oracle as O
    from ti5_isaacgym_b200.envs.t1.t1_robot import robot_constants
    from ti5_isaacgym_b200.sim.synthetic import (SyntheticTerrain, alloc_sim_tensors, fill_synthetic_state, synthetic_actions,
                                                 synthetic_height_field)
    cfg = make_cfg(num_envs, config)
    terrain = heights = None
    if config == 3:
        st = SyntheticTerrain(cfg.terrain, num_envs)
        terrain = SimpleNamespace(env_length=st.env_length, max_level=cfg.terrain.num_rows,
                                  origins=torch.from_numpy(st.env_origins).float().to(device))
        heights = synthetic_height_field(st.tot_rows, st.tot_cols, seed=7).to(device)
    C = O.make_consts(cfg, cfg.sim.dt, robot_constants(cfg), device=device, terrain=terrain)
    S = O.new_state(C, num_envs)
    gen = torch.Generator().manual_seed(seed)
    if terrain is not None:
        S.terrain_levels[:] = torch.randint(0, 6, (num_envs,), generator=gen).to(device)
        S.terrain_types[:] = torch.div(torch.arange(num_envs), num_envs / cfg.terrain.num_cols, rounding_mode="floor").long().to(device)
        S.env_origins[:] = terrain.origins[S.terrain_levels, S.terrain_types]
    sim = alloc_sim_tensors(num_envs, "cpu")
    fill_synthetic_state(sim, S.env_origins.cpu(), gen, base_contact_rate=CONTACT_RATE[config])
    sim = SimpleNamespace(**{k: v.to(device) for k, v in vars(sim).items()})
    S.episode_length_buf[:] = torch.randint(1, 2000, (num_envs,), generator=gen).to(device)
    S.gait_time[:, 1], S.gait_time[:, 2] = 900, 1500
    S.common_step_counter = COUNTER0[config]
    actions = synthetic_actions(num_envs, gen, "cpu").to(device)
    pools = {k: v.to(device) for k, v in O.draw_pools(C, num_envs, gen).items()}
    rew = torch.randn(ROLLOUT, num_envs, 1, generator=gen).to(device)
    val = torch.randn(ROLLOUT, num_envs, 1, generator=gen).to(device)
    done = (torch.rand(ROLLOUT, num_envs, 1, generator=gen) < 0.02).byte().to(device)
    last = torch.randn(num_envs, 1, generator=gen).to(device)
    sync = torch.cuda.synchronize if device != "cpu" else (lambda: None)
    count = [0]

    def one():
        O.step(C, S, sim, actions, pools, terrain=terrain, height_samples=heights)
        count[0] += 1
        if count[0] % ROLLOUT == 0:
            O.gae_returns(rew, val, done, last, GAMMA, LAM)

    with torch.inference_mode():
        for _ in range(warmup):
            one()
        sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        sync()
        dt = time.perf_counter() - t0
    return num_envs * steps / dt, dt / steps * 1e3


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 for N > 1: the CPU arm runs alone on rank 0 and takes every host core it may use
    try:
        torch.set_num_threads(host_threads())
    except RuntimeError:
        pass
    cores = torch.get_num_threads()
    # bounded sample: the whole run should end within ~2 minutes on the host cores (~30 us per env-step with all
    # threads), so large --steps shrink the number of envs stepped per step (never below 256, at most the workload's 8192)
    budget_env_steps = 3.0e6 if args.config == 2 else 1.5e6
    sample_envs = args.envs
    while sample_envs > 256 and sample_envs * (args.steps + args.warmup) > budget_env_steps:
        sample_envs //= 2
    value, ms = run_cpu_port(sample_envs, args.steps, args.warmup, config=args.config)
    line = {"impl": "reference", "metric": "env-steps/sec (step math + reward + obs)", "value": value,
            "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": workload_name(args.envs, args.config, 66), "baseline_config": args.config,
                                            "device": "host CPU", "cpu_model": cpu_model(), "physics": "no-op (fake gym)",
                                            "sample_envs": sample_envs},
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "cpu_model": cpu_model(),
                             "sample": f"{args.steps} steps x {sample_envs} envs of the oracle port (torch CPU, the reference's "
                                       "op chains, pinned bit-equal to the unmodified reference), all host threads"},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------

def dist_setup():
    import torch.distributed as dist
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
    return dist, world, rank, local, dev, group


def max_over_ranks(x, dev, world, dist):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def cuda_arm(args):
    dist, world, rank, local, dev, group = dist_setup()
    N, H = args.envs, args.frame_stack
    # The env step and its GAE shard with NO data-path collective: every rank normalises the advantages of its own
    # rollout, like the (single-process) reference does (rs:97-119).  The all-reduced statistics of a data-parallel PPO
    # run are part of `--ppo-iter` (BASELINE config 4), not of this metric: inside per-step event brackets a collective
    # measures how far eight free-running ranks have drifted apart since the last one (1216 M vs 1646 M env-steps/s at
    # 8 GPUs on the same box), not its own cost.
    loop = StepLoop(N, dev, args.config, H, rank, graph=os.environ.get("TI5_BENCH_GRAPH", "1") != "0",
                    materialize=args.materialize, group=None)
    env = loop.env
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(args.warmup, 3)
    # warm-up steps run exactly like the timed ones (flush, event pair, step): with a short warm-up (--warmup 5) the first
    # timed steps otherwise pay the host's first pass through the event / fill paths inside their brackets
    loop.timed(warm, flush)
    loop.warm_gae()
    # ---- device-timed region: K steps, L2 flushed (outside the event brackets) between steps ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    dev_ms, gae_in_region = loop.timed(args.steps, flush)
    barrier()
    dev_ms = max_over_ranks(dev_ms, dev, world, dist)
    value = world * N * args.steps / (dev_ms * 1e-3)

    # ---- per-launch duration of the kernel families (events on the launching stream) -------
    kt = kernel_times(env, loop.actions, steps=min(args.steps, 48))
    bracket_ms = bracket_overhead_ms(dev)

    # ---- end to end through env.step_host() with pinned host buffers ------------------------------------
    # pinned, device-mapped host buffers wired into the step's CUDA graph (LeggedRobot.enable_host_io): every
    # step_host() has the first kernel of the step read the actions from host memory, runs the step, has the observation
    # kernel store [rew | reset | time_outs] into host memory, and waits for the stream
    e2e = None
    if env._use_graph:
        h_act, h_out = env.enable_host_io()
        h_act.copy_(loop.actions.cpu())
        for _ in range(20):
            env.step_host()
        # wall-clock region: at least 2000 steps (0.1 s) — K = 240 steps are 12 ms, at the mercy of one scheduler hiccup —
        # with the clock sampler parked (one NVML query per second; the clocks were sampled every 5 ms through the
        # device-timed regions above): a query takes driver locks the launch path needs, and with the sampler running the
        # same loop measured anything between 139 and 168 M env-steps/s from run to run, against 166-168 M without it
        # (tools/e2e_probe.py)
        e2e_steps = max(args.steps, 2000)
        if sampler:
            sampler.interval = 1.0
            time.sleep(0.02)          # let a query in flight finish
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            loop.one(host=True)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0, dev, world, dist)
        e2e = {"value": world * N * e2e_steps / e2e_s, "unit": "env-steps/s", "steps": e2e_steps,
               "h2d_bytes_per_step": h_act.numel() * 4 * world,
               "d2h_bytes_per_step": h_out.numel() * world,
               "note": "env.step_host(): actions read from pinned host memory by the first kernel of the step, reward/reset/"
                       "time-out flags stored into pinned host memory by the last one (mapped pages, no copy-engine hop), "
                       "stream sync every step; L2 not flushed"}
    clocks = sampler.summary() if sampler else None      # sampled across the device-timed, per-phase and e2e regions

    # ---- BASELINE config 5: num_envs and frame-stack sweeps, every rank its own shard (same code path as above) ----
    sweep = None
    if not args.no_sweep and N == ENVS_PER_GPU and H == 66 and not args.materialize:
        del flush
        torch.cuda.empty_cache()
        points = [(1024, 66), (16384, 66), (65536, 66), (8192, 16), (8192, 100)] if args.config == 2 else [(65536, 66)]
        sweep = [sweep_point(n, h, dev, args.config, rank, world, dist, group) for n, h in points]

    if rank == 0:
        peak, peak_src = peaks()
        dom = max(kt, key=lambda k: kt[k]["share_ms"])
        bytes_per_launch = kt[dom]["bytes_per_launch"]
        ach = bytes_per_launch / (kt[dom]["ms_per_launch"] * 1e-3) / 1e9
        launches_per_step = env.launches_per_step
        fused = env._uses_fused_step()
        step_bytes = (BYTES_ENV_STEP_FUSED if fused else BYTES_ENV_STEP) + (BYTES_HEIGHTS + 12 * 187 if args.config == 3 else 0)
        cpu = torch_gpu = None
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(host_threads())
            v, ms = run_cpu_port(4096, 30, 3, config=args.config)
            cores = torch.get_num_threads()
            torch.set_num_threads(1)
            v1, ms1 = run_cpu_port(1024, 12, 2, config=args.config)
            torch.set_num_threads(cores)
            cpu = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "cpu_model": cpu_model(),
                   "sample": f"30 steps x 4096 envs (BASELINE config 1 size) of the oracle port, torch CPU, all {cores} host "
                             f"threads, {ms:.1f} ms/step",
                   "single_thread": {"value": v1, "cores": 1,
                                     "sample": f"12 steps x 1024 envs, torch.set_num_threads(1), {ms1:.1f} ms/step"}}
            # the reference's eager-torch algorithm on THIS GPU (the north_star's 20x denominator); context only
            v, ms = run_cpu_port(N, 24, 3, device=dev, config=args.config)
            torch_gpu = {"value": v, "unit": "env-steps/s", "ms_per_step": ms,
                         "sample": f"24 steps x {N} envs of the oracle port in eager torch on the same B200"}
        rollout = None
        if world == 1 and not args.no_rollout and not args.materialize and args.config == 2:
            rollout = rollout_bench(env, loop.gen, loop.actions)
        line = {
            "metric": "env-steps/sec (step math + reward + obs)", "value": value, "unit": "env-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(N, args.config, H), "baseline_config": args.config,
                       "envs_per_gpu": N, "frame_stack": H, "obs": "materialised" if args.materialize else "ring view",
                       "rng": "in-kernel Philox4x32-10", "physics": "no-op (synthetic state, SURVEY 8d)",
                       "gae_calls_in_timed_region": gae_in_region,
                       "parallelism": f"{world} x env shards, no data-path collective (advantages normalised per rank)",
                       "l2": "flushed between timed steps (256 MiB write outside the event brackets)",
                       "launch": (f"one CUDA graph per step ({launches_per_step} kernels" if env._use_graph
                                  else f"direct launches ({launches_per_step} kernels")
                                 + (", programmatic dependent launches)" if env._chain_launches else ")")},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps + 2 * gae_in_region,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": ncu_traffic(dom) if N == ENVS_PER_GPU and args.config == 2 else None,
                         "peak_source": peak_src, "bytes_per_launch": bytes_per_launch,
                         "ms_per_launch": kt[dom]["ms_per_launch"],
                         "empty_bracket_ms": bracket_ms,
                         "whole_step": {"bytes": step_bytes * N, "achieved": step_bytes * N * args.steps / (dev_ms * 1e-3) / 1e9,
                                        "frac": step_bytes * N * args.steps / (dev_ms * 1e-3) / 1e9 / peak,
                                        "survey_8d_bytes": BYTES_ENV_STEP * N,
                                        "survey_8d_frac": BYTES_ENV_STEP * N * args.steps / (dev_ms * 1e-3) / 1e9 / peak},
                         "kernels": kt,
                         "how": "CUDA events on the launching stream around each kernel family of the step (L2 flushed before "
                                "each step); per-launch = family time / launches in the family; algorithmic bytes = DESIGN.md "
                                "section 4 per-env figure x envs.  At 8192 envs every kernel is latency-bound (13 us of HBM time "
                                "per step): see `sweep` for the sizes where bandwidth is the bound.  `empty_bracket_ms` = what the same event "
                                "bracket reads around one empty kernel (launch / event latency inside every per-family time)"},
            "cpu_baseline": cpu,
            "reference_torch_gpu": torch_gpu,
            "rollout_storage": rollout,
            "sweep": sweep,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sweep_point(N, H, dev, config, rank, world, dist, group, steps=48, warmup=12):
    """One point of BASELINE config 5 (num_envs / frame_stack sweep): the same device-timed step (L2 flushed before every
    step, GAE every 24th) and per-family timing as the main line, on a fresh env of N envs per GPU; under torchrun every
    rank steps its own shard and the time is the max over ranks."""
    loop = StepLoop(N, dev, config, H, rank, group=None, seed=4321)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        loop.one()
    loop.warm_gae()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms, _ = loop.timed(steps, flush)
    ms = max_over_ranks(ms, dev, world, dist) / steps
    kt = kernel_times(loop.env, loop.actions, steps=24)
    peak, _ = peaks()
    step_bytes = BYTES_ENV_STEP_FUSED if loop.env._uses_fused_step() else BYTES_ENV_STEP     # the frames written do not depend on H
    out = {"envs_per_gpu": N, "frame_stack": H, "n_gpus": world, "value": world * N / (ms * 1e-3), "unit": "env-steps/s",
           "ms_per_step": ms,
           "kernels": {k: {"ms_per_launch": v["ms_per_launch"], "achieved": v["achieved_GBps"], "frac": v["achieved_GBps"] / peak}
                       for k, v in kt.items()},
           "whole_step": {"bytes": step_bytes * N, "achieved": step_bytes * N / (ms * 1e-3) / 1e9,
                          "frac": step_bytes * N / (ms * 1e-3) / 1e9 / peak,
                          "survey_8d_frac": BYTES_ENV_STEP * N / (ms * 1e-3) / 1e9 / peak}}
    del loop, flush
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


def bracket_overhead_ms(dev, reps=48):
    """What a CUDA-event bracket around ONE (almost) empty kernel reads, after the same L2 flush: the share of every
    per-family time below that is launch / event latency, not kernel execution (calibration only: a torch fill of one
    element, none of this library's kernels)."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    x = torch.zeros(1, device=dev)
    ts = []
    for i in range(reps + 4):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        x.fill_(1.0)
        b.record()
        if i >= 4:
            ts.append((a, b))
    torch.cuda.synchronize()
    del flush
    return statistics.mean(a.elapsed_time(b) for a, b in ts)


def kernel_times(env, actions, steps):
    """Average device time of the kernel families of a step: CUDA events on the launching stream around each family
    (direct launches, L2 flushed before every step).  A single whole-step graph cannot be bracketed inside, and the
    event records cut the programmatic launch chain at the family boundaries, so the times add up to more than the
    whole-step time.  Per-launch time = family time / launches in the family."""
    phases = env.phase_launchers()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=env.device)
    env._actions_in.copy_(actions)
    acc = {name: [] for name, _, _ in phases}
    marks = []
    for i in range(steps + 4):
        flush.fill_(i & 0xFF)
        e = [ev() for _ in range(len(phases) + 1)]
        e[0].record()
        for k, (_, fn, _) in enumerate(phases):
            fn()
            e[k + 1].record()
        env._finish_step()
        if i >= 4:
            marks.append(e)
    torch.cuda.synchronize()
    for e in marks:
        for k, (name, _, _) in enumerate(phases):
            acc[name].append(e[k].elapsed_time(e[k + 1]))
    out = {}
    for name, _, n in phases:
        phase = statistics.mean(acc[name])
        extra = env._params.priv_frame - 73 if name == "reset_observe" else 0       # measured heights ride in the critic frame
        nbytes = (KERNEL_BYTES[name] + 4 * 3 * extra) * env.num_envs
        out[name] = {"phase_ms": phase, "launches": n, "ms_per_launch": phase / n, "share_ms": phase,
                     "bytes_per_launch": nbytes, "achieved_GBps": nbytes / (phase / n * 1e-3) / 1e9}
    del flush
    return out


# ---------------------------------------------------------------------------------------------
# BASELINE config 4 as written: envs sharded 8192/GPU WITH the PPO gradient all-reduce — a whole PPO iteration
# ---------------------------------------------------------------------------------------------

def ppo_iter_arm(args):
    """One PPO iteration as `DHOnPolicyRunner.learn` runs it (dh_on_policy_runner.py:125-187) on every rank's shard:
    24 x [policy forward -> env.step -> ti5_store_transition], GAE with the 3-double all-reduce of the advantage
    statistics, then 2 epochs x 4 mini-batches (t1_cfg:455-468) of: ti5_gather_minibatch -> forward/backward of a
    stand-in actor-critic with ActorCriticDH's parameter count -> ONE flat-bucket gradient all-reduce
    (FlatGradAllReduce) -> clip -> Adam.  The actor-critic itself is out of scope (it stays plain PyTorch); the point is
    where the collective sits and what it costs next to the env steps.  Device-timed, max over ranks."""
    dist, world, rank, local, dev, group = dist_setup()
    from ti5_isaacgym_b200.algo.rollout_storage import FrameLogRolloutStorage, RolloutStorage
    from ti5_isaacgym_b200.distributed import FlatGradAllReduce, all_reduce_mean
    N, T, EPOCHS, NMB = args.envs, ROLLOUT, 2, 4
    loop = StepLoop(N, dev, 2, 66, rank, group=group)
    env = loop.env
    storage = FrameLogRolloutStorage(env, T, group=group)
    torch.manual_seed(0)
    P = env.num_privileged_obs

    class ActorCritic(torch.nn.Module):          # stand-in with ActorCriticDH's size (~0.86 M parameters, actor_critic_dh.py:31-117)
        def __init__(self):
            super().__init__()
            mlp = lambda i, o: torch.nn.Sequential(torch.nn.Linear(i, 512), torch.nn.ELU(), torch.nn.Linear(512, 256), torch.nn.ELU(),
                                                   torch.nn.Linear(256, 128), torch.nn.ELU(), torch.nn.Linear(128, o))
            self.hist = torch.nn.Conv1d(66, 32, 6, stride=3)          # long-history encoder over (N, 66, 47)
            self.actor, self.critic = mlp(235 + 32 * 14, 12), mlp(P, 1)
            self.std = torch.nn.Parameter(torch.ones(12))

        def forward(self, obs, critic_obs):
            h = self.hist(obs.view(-1, 66, 47)).flatten(1)
            return self.actor(torch.cat((obs[..., -235:], h), -1)), self.critic(critic_obs)

    net = ActorCritic().to(dev)
    n_params = sum(p.numel() for p in net.parameters())
    opt = torch.optim.Adam(net.parameters(), lr=1e-5)
    sync = FlatGradAllReduce(net, group)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    phases = {"collect": [], "gae": [], "update": [], "allreduce": [], "iteration": []}

    def iteration(record):
        e = [ev() for _ in range(4)]
        e[0].record()
        storage.clear()
        obs, priv = env.get_observations(), env.get_privileged_observations()
        with torch.inference_mode():
            for _ in range(T):
                mu, v = net(obs, priv)
                act = mu + net.std * torch.randn_like(mu)
                tr = RolloutStorage.Transition()
                tr.actions, tr.values, tr.action_mean, tr.action_sigma = act, v, mu, net.std.expand_as(mu).contiguous()
                tr.actions_log_prob = -0.5 * ((act - mu) / net.std).square().sum(-1)
                tr.observations, tr.critic_observations = obs, priv
                obs, priv, rew, dones, infos = env.step(act)
                storage.store_step(tr, rew, dones, infos["time_outs"], GAMMA)
            e[1].record()
            storage.compute_returns(net(obs, priv)[1], GAMMA, LAM)       # advantage statistics all-reduced over the ranks
        e[2].record()
        ar = []
        for _ in range(EPOCHS):
            perm = torch.randperm(T * N, device=dev)
            for i in range(NMB):
                b = storage.gather(perm[i * (T * N // NMB):(i + 1) * (T * N // NMB)])
                mu, v = net(b["obs"], b["critic_obs"])
                logp = -0.5 * ((b["actions"] - mu) / net.std).square().sum(-1, keepdim=True)
                ratio = torch.exp(logp - b["actions_log_prob"])
                adv = b["advantages"]
                loss = torch.max(-adv * ratio, -adv * ratio.clamp(0.8, 1.2)).mean() + (b["returns"] - v).square().mean()
                all_reduce_mean((logp - b["actions_log_prob"]).mean().detach().reshape(1), group)      # KL, dh_ppo.py:139-151
                opt.zero_grad(set_to_none=False)
                loss.backward()
                a, z = ev(), ev()
                a.record(); sync.reduce(); z.record()          # dh_ppo.py:180-181: between backward and the clip
                ar.append((a, z))
                torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
                opt.step()
        e[3].record()
        if record:
            torch.cuda.synchronize()
            phases["collect"].append(e[0].elapsed_time(e[1])); phases["gae"].append(e[1].elapsed_time(e[2]))
            phases["update"].append(e[2].elapsed_time(e[3])); phases["iteration"].append(e[0].elapsed_time(e[3]))
            phases["allreduce"].append(sum(a.elapsed_time(z) for a, z in ar))

    iters = max(2, min(args.steps, 6))
    for _ in range(2):
        iteration(False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for _ in range(iters):
        iteration(True)
    mean = {k: max_over_ranks(statistics.mean(v), dev, world, dist) for k, v in phases.items()}
    if rank == 0:
        it_ms = mean["iteration"]
        line = {"metric": "PPO iteration (24 env steps + GAE + 8 mini-batch updates) env-steps/sec", "mode": "ppo-iter",
                "value": world * N * T / (it_ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": iters, "warmup": 2,
                "ms_per_step": it_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"BASELINE config 4: {world * N} envs sharded {N}/GPU over {world} GPU(s), whole PPO iteration "
                                       "with the gradient all-reduce", "envs_per_gpu": N, "rollout": T, "epochs": EPOCHS,
                           "mini_batches": NMB, "stand_in_parameters": n_params},
                "iteration_ms": it_ms, "collect_ms": mean["collect"], "gae_ms": mean["gae"], "update_ms": mean["update"],
                "grad_allreduce": {"calls_per_iteration": EPOCHS * NMB, "us_per_call": 1e3 * mean["allreduce"] / (EPOCHS * NMB),
                                   "bytes_per_call": 4 * n_params, "share_of_iteration": mean["allreduce"] / it_ms,
                                   "how": "CUDA events around FlatGradAllReduce.reduce() (bucket copy + NCCL all-reduce, AVG)"},
                "env_steps_share_of_iteration": mean["collect"] / it_ms}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=240)
    ap.add_argument("--warmup", type=int, default=24)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU (BASELINE metric: 8192)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3],
                    help="BASELINE.json config: 2 = plane, 3 = trimesh + heights + pushes + 5 %% resets")
    ap.add_argument("--frame-stack", type=int, default=66, help="observation history length H (BASELINE config 5 sweeps it)")
    ap.add_argument("--materialize", action="store_true", help="also write contiguous (N,3102)/(N,219) observations")
    ap.add_argument("--ppo-iter", action="store_true", help="BASELINE config 4: a whole PPO iteration with the gradient all-reduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rollout", action="store_true", help="skip the rollout-storage measurement (SURVEY 8f rows 1-2)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the num_envs / frame_stack sweep points (BASELINE config 5)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    elif args.ppo_iter:
        ppo_iter_arm(args)
    else:
        cuda_arm(args)


if __name__ == "__main__":
    main()
