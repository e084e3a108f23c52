"""Host-side enqueue cost of one update vs its device time (is the launch loop the bottleneck?)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import deep_rl_b200 as drl
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tr = drl.PPOTrainer(drl.PPOConfig(num_envs=N, num_steps=128, total_timesteps=N * 128 * 1000))
for _ in range(5):
    tr.update(1000)
torch.cuda.synchronize()
K = 50
t0 = time.perf_counter()
for _ in range(K):
    tr.update(1000)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"N={N}: host enqueue {1e3 * (t1 - t0) / K:.3f} ms/update, enqueue+drain {1e3 * (t2 - t0) / K:.3f} ms/update")
