"""End-to-end rate of env.step_host() over many steps (the bench's e2e leg runs K = 240 steps = 13 ms of wall clock, which
is at the mercy of host jitter), next to the warm device-timed step in one event bracket.
    TI5_LIB=... python tools/e2e_probe.py [envs] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from bench import StepLoop

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
loop = StepLoop(N, "cuda:0", 2, 66, rank=0)
env = loop.env
for _ in range(24):
    loop.one()
h_act, h_out = env.enable_host_io()
h_act.copy_(loop.actions.cpu())
for _ in range(20):
    env.step_host()
res = []
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(K):
        loop.one(host=True)
    dt = time.perf_counter() - t0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        loop.one()
    b.record()
    torch.cuda.synchronize()
    res.append((dt / K * 1e6, a.elapsed_time(b) / K * 1e3))
# the same host-I/O graph back to back inside one event bracket (no host wait): device time with the PCIe reads / writes
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(K):
    env._graph_host.replay()
b.record()
torch.cuda.synchronize()
print("host-I/O graph back to back: %.2f us/step (device)" % (a.elapsed_time(b) / K * 1e3))
# floor of a launch + wait round trip: an (almost) empty graph replayed and waited for
x = torch.zeros(1, device="cuda:0")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    x.add_(1.0)
for _ in range(20):
    g.replay(); torch.cuda.current_stream().synchronize()
t0 = time.perf_counter()
for _ in range(K):
    g.replay(); torch.cuda.current_stream().synchronize()
print("empty graph replay + stream sync: %.2f us per round trip" % ((time.perf_counter() - t0) / K * 1e6))
print(os.path.basename(os.environ.get("TI5_LIB", "product")), N, "e2e us/step", [round(r[0], 2) for r in res],
      "warm device us/step (one bracket)", [round(r[1], 2) for r in res])
