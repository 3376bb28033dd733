import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the probe sites are compiled in with -DTI5_PROBES only: an instrumented library next to the product one
PROBE_LIB = os.path.join(ROOT, "tools", "exp", "libti5_probes.so")
if not os.path.exists(PROBE_LIB) or os.environ.get("REBUILD_PROBES") == "1":
    from ti5_isaacgym_b200.build import build_variant
    os.makedirs(os.path.dirname(PROBE_LIB), exist_ok=True)
    build_variant(PROBE_LIB, ["-DTI5_PROBES"])
os.environ["TI5_LIB"] = PROBE_LIB
import torch, sys
from bench import make_cfg
from ti5_isaacgym_b200.envs import T1DHStandEnv
from ti5_isaacgym_b200.sim.synthetic import SimParams, fill_synthetic_state, synthetic_actions
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
CONFIG3 = len(sys.argv) > 2 and sys.argv[2] == "config3"        # trimesh + measured heights + pushes + 5 % resets per step
cfg = make_cfg(N, 3 if CONFIG3 else 2)
env = T1DHStandEnv(cfg, SimParams(dt=cfg.sim.dt), 1, 'cuda:0', True, use_cuda_graph=True, materialize_obs=False)
gen = torch.Generator(device='cuda').manual_seed(1)
fill_synthetic_state(env.gym.tensors, env.env_origins, gen, base_contact_rate=0.05 if CONFIG3 else 0.01)
env.reset()
env.episode_length_buf = torch.randint(1, 2000, (N,), generator=gen, device='cuda')
env._debug_ts = torch.zeros(3, 4096, 8, dtype=torch.int64, device='cuda')
env._bind_buffers(); env._drop_graphs()
act = synthetic_actions(N, gen, 'cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
NOFLUSH = os.environ.get('NOFLUSH') == '1'
CLEAN = os.environ.get('CLEANFLUSH') == '1'          # write the flush buffer, then read a second one: the L2 ends up full of CLEAN lines
flush2 = torch.zeros(256 << 18, dtype=torch.int32, device='cuda') if CLEAN else None
for i in range(20):
    if not NOFLUSH: flush.fill_(i)
    if CLEAN: sink = flush2.sum()
    env.step(act)
torch.cuda.synchronize()
ts = env._debug_ts.cpu().double()
for kern, name, nprobe in ((0, 'post_physics / fused_step', 8), (2, 'substep workers', 2), (1, 'reset_observe', 8)):
    t = ts[kern]
    used = t[:, 0] > 0
    t = t[used]
    if not len(t):
        continue
    t0 = ts[0][:, 0][ts[0][:, 0] > 0].min()          # everything relative to the first CTA of the step's first per-env kernel
    print(name, 'CTAs', int(used.sum()))
    for k in range(nprobe):
        col = t[:, k]; col = col[col > 0]
        if len(col): print('  probe %d: n=%4d  min %7.2f  median %7.2f  max %7.2f us' % (k, len(col), (col.min()-t0)/1e3, (col.median()-t0)/1e3, (col.max()-t0)/1e3))
# reset_observe by CTA class: env CTAs holding a re-spawned env, the other env CTAs, helper CTAs
counts = env._block_counts.cpu()[:(N + env._params.env_block - 1) // env._params.env_block]
t = ts[1]
nb = len(counts)
t0 = ts[0][:, 0][ts[0][:, 0] > 0].min()
for label, rows in (("env CTAs with a reset", (counts > 0).nonzero().flatten()), ("env CTAs without", (counts == 0).nonzero().flatten()),
                    ("helper CTAs", torch.arange(nb, int((t[:, 0] > 0).sum())))):
    print(label, len(rows))
    for k in (0, 6, 1, 2, 3, 7, 4, 5):
        col = t[rows, k]; col = col[col > 0]
        if len(col): print('  probe %d: median %7.2f  max %7.2f us' % (k, (col.median()-t0)/1e3, (col.max()-t0)/1e3))
b1_end = ts[0][:, 5].max(); b2_start = ts[1][:, 0][ts[1][:, 0] > 0].min()
print('gap B1 end -> B2 start: %.2f us' % ((b2_start - b1_end) / 1e3))
# per-CTA stretches (probe k minus the same CTA's probe 0): meaningful on grids of several waves, where the absolute
# times above mix the waves
for kern, name in ((0, 'post_physics / fused_step'), (1, 'reset_observe')):
    t = ts[kern]
    t = t[t[:, 0] > 0]
    if kern == 1: t = t[:nb]
    print(name, 'per-CTA time since its probe 0 (env CTAs)')
    for k in range(1, 8):
        ok = t[:, k] > 0
        if ok.any():
            d = (t[ok, k] - t[ok, 0]) / 1e3
            print('  probe %d: n=%4d  median %7.2f  p90 %7.2f  max %7.2f us' % (k, int(ok.sum()), d.median(), d.quantile(0.9), d.max()))
# builder probes of reset_observe (kernel row 2, slots 2-7), relative to the CTA's probe 0 (behind the grid wait)
t2, t1 = ts[2][:nb], ts[1][:nb]
names = {7: 'counts folded', 2: 'obs A: loads issued, second command pass done', 3: 'obs A done', 4: 'obs B done', 5: 'priv A done', 6: 'priv B done'}
print('reset_observe builders, per-CTA time since probe 0')
for k in (7, 2, 3, 4, 5, 6):
    ok = (t2[:, k] > 0) & (t1[:, 0] > 0)
    if ok.any():
        d = (t2[ok, k] - t1[ok, 0]) / 1e3
        print('  %-48s median %6.2f  p90 %6.2f  max %6.2f us' % (names[k], d.median(), d.quantile(0.9), d.max()))
