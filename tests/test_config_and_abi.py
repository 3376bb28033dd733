"""CPU tests: configuration trees, the C-ABI library (symbols, struct layout, argument checks) and
the cfg -> Ti5Params host logic.  No compute call is made (there is no GPU here)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from helpers import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_trees_match_reference_dump():
    from ti5_isaacgym_b200.envs import DHT1StandCfg, DHT1StandCfgPPO, LeggedRobotCfg, LeggedRobotCfgPPO
    from ti5_isaacgym_b200.utils.helpers import class_to_dict
    want = json.load(open(os.path.join(GOLDEN, "reference_cfg_tree.json")))
    for cls in (DHT1StandCfg, DHT1StandCfgPPO, LeggedRobotCfg, LeggedRobotCfgPPO):
        got = json.loads(json.dumps(class_to_dict(cls())))
        assert got == want[cls.__name__], cls.__name__


def test_reward_terms_are_alphabetical():
    from ti5_isaacgym_b200.envs import DHT1StandCfg
    from ti5_isaacgym_b200.envs.base.step_params import TERM_NAMES, reward_scales
    names = list(reward_scales(DHT1StandCfg(), 0.01))
    assert names == sorted(names) and len(names) == 24
    assert list(TERM_NAMES) == sorted(TERM_NAMES)
    assert set(names) <= set(TERM_NAMES)


def test_history_length_variants():
    from ti5_isaacgym_b200.envs import make_t1_cfg, make_t1_cfg_ppo
    for H in (1, 10, 66, 100):
        cfg = make_t1_cfg(frame_stack=H)
        assert cfg.env.num_observations == 47 * H
        assert make_t1_cfg_ppo(cfg).policy.in_channels == H
    assert make_t1_cfg_ppo(make_t1_cfg()).algorithm.lin_vel_idx == 73 * 2 + 53


def test_library_exports_every_declared_symbol():
    from ti5_isaacgym_b200 import _lib
    lib = _lib.load_library()
    header = open(os.path.join(ROOT, "include", "ti5_step.h")).read()
    declared = set(re.findall(r"\b(ti5_\w+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ti5_version() == _lib.CONSTS["TI5_ABI_VERSION"]
    sizes = (ctypes.c_int32 * 4)()
    assert lib.ti5_struct_sizes(sizes) == 0
    assert list(sizes) == [ctypes.sizeof(s) for s in (_lib.Ti5Params, _lib.Ti5Buffers, _lib.Ti5Rng, _lib.Ti5Globals)]
    sizes = (ctypes.c_int32 * 3)()
    assert lib.ti5_rollout_struct_sizes(sizes) == 0
    assert list(sizes) == [ctypes.sizeof(s) for s in (_lib.Ti5Rollout, _lib.Ti5Transition, _lib.Ti5Batch)]


def test_reward_name_table_matches_binding():
    from ti5_isaacgym_b200 import _lib
    from ti5_isaacgym_b200.envs.base.step_params import TERM_NAMES
    assert _lib.reward_names() == list(TERM_NAMES)
    assert _lib.load_library().ti5_reward_name(99) is None


def test_invalid_arguments_are_reported_not_thrown():
    from ti5_isaacgym_b200 import _lib
    lib = _lib.load_library()
    p = _lib.Ti5Params()          # num_envs = 0
    b = _lib.Ti5Buffers()
    rc = lib.ti5_begin_step(ctypes.byref(p), ctypes.byref(b), None, None)
    assert rc == _lib.CONSTS["TI5_EINVAL"]
    assert b"ti5_begin_step" in lib.ti5_last_error()
    assert lib.ti5_post_physics(None, None, None, 0, None) == _lib.CONSTS["TI5_EINVAL"]
    assert lib.ti5_gae(None, None, None, None, None, None, 24, 8, 0.9, 0.9, None, None, None) == _lib.CONSTS["TI5_EINVAL"]
    with pytest.raises(_lib.Ti5Error):
        _lib.check(rc)


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from ti5_isaacgym_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(_lib.Ti5Error, match="no CPU"):
        _lib.load_library(str(tmp_path / "absent.so"))


def test_env_refuses_cpu_device():
    from ti5_isaacgym_b200.envs import DHT1StandCfg, T1DHStandEnv
    from ti5_isaacgym_b200.sim.synthetic import SimParams
    cfg = DHT1StandCfg()
    cfg.env.num_envs = 4
    cfg.terrain.mesh_type = "plane"
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        T1DHStandEnv(cfg, SimParams(), 1, "cpu", True)


def test_gae_refuses_cpu_tensors():
    import torch
    from ti5_isaacgym_b200 import _lib
    from ti5_isaacgym_b200.algo.rollout_storage import RolloutStorage
    st = RolloutStorage(8, 4, [3], [3], [2])
    with pytest.raises((_lib.Ti5Error, RuntimeError)):
        st.compute_returns(torch.zeros(8, 1), 0.99, 0.9)


def test_params_match_oracle_constants():
    """Every derived scalar the kernels use equals what the oracle (pinned to the reference) derives."""
    import torch
    from oracle import t1_oracle as O
    from ti5_isaacgym_b200.envs import DHT1StandCfg
    from ti5_isaacgym_b200.envs.base.step_params import TERM_NAMES, build_params
    from ti5_isaacgym_b200.envs.t1.t1_robot import robot_constants
    cfg = DHT1StandCfg()
    cfg.terrain.mesh_type = "plane"
    robot = robot_constants(cfg)
    C = O.make_consts(cfg, cfg.sim.dt, robot)
    p = build_params(cfg, cfg.sim.dt, robot)
    f32 = lambda x: float(np.float32(x))
    assert p.dt == f32(C.dt) and p.max_episode_length == int(C.max_episode_length) == 2400
    assert p.ext_force_interval == int(C.ext_force_interval) and p.push_interval == int(C.push_interval)
    assert list(p.torque_limits) == C.torque_limits.tolist()
    assert list(p.dof_vel_limits) == C.dof_vel_limits.tolist()
    assert list(p.default_dof_pos) == C.default_dof_pos[0].tolist()
    assert list(p.p_gains) == C.p_gains.tolist() and list(p.d_gains) == C.d_gains.tolist()
    assert list(p.noise_vec)[:47] == C.noise_scale_vec.tolist()
    for name, s in C.reward_scales.items():
        t = TERM_NAMES.index(name)
        assert p.reward_scale[t] == f32(s) and (p.term_mask >> t) & 1
    assert bin(p.term_mask).count("1") == len(C.reward_scales)
    assert [p.feet[0], p.feet[1], p.knees[0], p.knees[1], p.term_body, p.pen_body] == [6, 12, 4, 10, 0, 0]
    assert p.lag_len == 31 and p.dof_lag_len == 31 and p.imu_lag_len == 11
    assert p.torque_multi_w == f32(1.2 - 0.8) and p.torque_multi_lo == f32(0.8)
    assert [p.add_duration[i] for i in range(p.n_add_dur)] == [d / C.dt for d in cfg.domain_rand.add_duration]


def test_env_block_choice_fills_the_sms():
    from ti5_isaacgym_b200.envs.base.step_params import pick_env_block
    assert pick_env_block(8192) == 32 and pick_env_block(65536) == 64 and pick_env_block(20000) == 64


def test_unsupported_options_raise():
    from ti5_isaacgym_b200.envs import DHT1StandCfg
    from ti5_isaacgym_b200.envs.base.step_params import build_params
    from ti5_isaacgym_b200.envs.t1.t1_robot import robot_constants
    from ti5_isaacgym_b200._lib import CONSTS
    cfg = DHT1StandCfg()
    cfg.domain_rand.randomize_joint_friction = True     # one multiplier per env: supported (lr:762-763)
    p = build_params(cfg, 0.001, robot_constants(cfg))
    assert p.flags2 == CONSTS["TI5_F2_RAND_JOINT_FRICTION"] and abs(p.joint_friction_lo - 0.01) < 1e-7
    cfg.domain_rand.randomize_joint_friction_each_joint = True   # per-joint ranges exist for ten of the twelve joints only
    with pytest.raises(NotImplementedError):
        build_params(cfg, 0.001, robot_constants(cfg))
    cfg = DHT1StandCfg()                                # the lag options t1_cfg marks "always False" (t1_cfg:290-312)
    cfg.domain_rand.randomize_lag_timesteps_perstep = True
    cfg.domain_rand.randomize_imu_lag_timesteps_perstep = True
    p = build_params(cfg, 0.001, robot_constants(cfg))
    assert p.flags & CONSTS["TI5_F_LAG_PERSTEP"] and p.flags & CONSTS["TI5_F_IMU_LAG_PERSTEP"]
    assert not p.flags & CONSTS["TI5_F_DOF_LAG_PERSTEP"] and not p.flags & CONSTS["TI5_F_POS_VEL_LAG"]
    cfg = DHT1StandCfg()
    cfg.domain_rand.add_dof_pos_vel_lag = True          # shows only with the common joint-state lag off (t1:407, 416)
    assert not build_params(cfg, 0.001, robot_constants(cfg)).flags & CONSTS["TI5_F_POS_VEL_LAG"]
    cfg.domain_rand.add_dof_lag = False
    p = build_params(cfg, 0.001, robot_constants(cfg))
    assert p.flags & CONSTS["TI5_F_POS_VEL_LAG"] and p.flags & CONSTS["TI5_F_ADD_DOF_LAG"] and p.dof_lag_len == 26
    cfg = DHT1StandCfg()
    cfg.commands.heading_command = True                 # supported since round 2 (t1:141-176, 185-188)
    p = build_params(cfg, 0.001, robot_constants(cfg))
    from ti5_isaacgym_b200._lib import CONSTS
    assert p.flags & CONSTS["TI5_F_HEADING_COMMAND"] and abs(p.heading_w - 6.28) < 1e-6 and abs(p.heading_lo + 3.14) < 1e-6
    assert not p.flags & CONSTS["TI5_F_NO_SW_SWITCH"]
    cfg.commands.sw_switch = False                      # t1:89-90
    assert build_params(cfg, 0.001, robot_constants(cfg)).flags & CONSTS["TI5_F_NO_SW_SWITCH"]
    cfg = DHT1StandCfg()
    cfg.rewards.scales.dof_vel_limits = -1.0
    with pytest.raises(NotImplementedError):
        build_params(cfg, 0.001, robot_constants(cfg))


def test_env_block_follows_the_grid_size():
    """step_params.pick_env_block: 32 while the grid is at most two CTAs per SM (the BASELINE size), 64 beyond."""
    from ti5_isaacgym_b200.envs.base.step_params import pick_env_block
    assert [pick_env_block(n) for n in (1, 1024, 8192, 9472, 9473, 12288, 16384, 65536)] == [32, 32, 32, 32, 64, 64, 64, 64]


def test_vec_env_contract_reports_what_is_missing():
    import torch
    from types import SimpleNamespace
    from ti5_isaacgym_b200.algo.vec_env import VecEnv, check_vec_env
    ok = SimpleNamespace(num_envs=4, num_obs=6, num_short_obs=2, num_privileged_obs=None, num_actions=3, max_episode_length=2400.0,
                         obs_buf=torch.zeros(4, 6), privileged_obs_buf=None, rew_buf=torch.zeros(4), reset_buf=torch.ones(4, dtype=torch.bool),
                         episode_length_buf=torch.zeros(4, dtype=torch.int64), extras={}, device="cpu",
                         step=lambda a: None, reset=lambda: None, get_observations=lambda: None,
                         get_privileged_observations=lambda: None)
    assert check_vec_env(ok) == [] and isinstance(ok, VecEnv)
    ok.rew_buf = torch.zeros(5, dtype=torch.float64)
    del ok.get_observations
    faults = check_vec_env(ok)
    assert any("get_observations" in f for f in faults) and any("rew_buf should be torch.float32" in f for f in faults)
    assert any("rew_buf should have shape (4,)" in f for f in faults) and not isinstance(ok, VecEnv)
