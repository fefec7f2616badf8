"""Host-side cost of one `vmap_step` call through the Python mirror (B = 5: the launch-latency regime)."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import exciting_environments_b200 as ex  # noqa: E402

B, N = 5, 2000
env = ex.Pendulum(batch_size=B, tau=2e-2)
_, state = env.vmap_reset()
action = torch.zeros((B, 1), device="cuda")
for _ in range(200):
    _, state = env.vmap_step(state, action)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(N):
    _, state = env.vmap_step(state, action)
torch.cuda.synchronize()
print(f"vmap_step: {(time.perf_counter() - t0) / N * 1e6:.1f} us/call")
gym = ex.GymWrapper(env)
gym.reset()
for _ in range(200):
    gym.step(action)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(N):
    gym.step(action)
torch.cuda.synchronize()
print(f"GymWrapper.step: {(time.perf_counter() - t0) / N * 1e6:.1f} us/call")
if len(sys.argv) > 1:
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(N):
        _, state = env.vmap_step(state, action)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
