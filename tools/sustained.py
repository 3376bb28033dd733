"""How much of the device-timed step is launch / event latency around the step's graph, and how much is cold code:
  (a) per-step CUDA events, L2 flushed (256 MiB write) before every step  — what bench.py calls `isolated`;
  (b) K steps back to back inside ONE event bracket, nothing in between    — warm L2, launches pipelined;
  (c) K steps inside one bracket, round-robin over M independent env instances whose combined per-step working set
      exceeds the L2 several times: every step's data is cold, the code is warm, launches are pipelined.
    python tools/sustained.py [envs] [instances]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from bench import StepLoop

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
M = int(sys.argv[2]) if len(sys.argv) > 2 else 6
K = 240
dev = "cuda:0"
loops = [StepLoop(N, dev, 2, 66, rank=0, seed=1234 + 17 * i) for i in range(M)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
for lp in loops:
    for _ in range(24):
        lp.one()
torch.cuda.synchronize()
for rep in range(2):
    ms, _ = loops[0].timed(K, flush)
    print("(a) isolated, flushed, per-step events: %.2f us/step" % (ms / K * 1e3))
    a, b = ev(), ev()
    a.record()
    for _ in range(K):
        loops[0].one()
    b.record()
    torch.cuda.synchronize()
    print("(b) one bracket, same instance, warm:    %.2f us/step" % (a.elapsed_time(b) / K * 1e3))
    a, b = ev(), ev()
    a.record()
    for i in range(K):
        loops[i % M].one()
    b.record()
    torch.cuda.synchronize()
    print("(c) one bracket, %d instances round-robin: %.2f us/step" % (M, a.elapsed_time(b) / K * 1e3))
    ts = []
    for i in range(K):
        s, e = ev(), ev()
        s.record(); loops[i % M].one(); e.record()
        ts.append((s, e))
    torch.cuda.synchronize()
    print("(d) per-step events, %d instances round-robin, no flush: %.2f us/step" % (M, sum(s.elapsed_time(e) for s, e in ts) / K * 1e3))
