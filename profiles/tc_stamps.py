"""Cycle-stamp timeline of ppo_grad_tc_kernel (CTA 0).  Needs a library built with -DDRL_TC_STAMPS:
   DRL_EXTRA_NVCC_FLAGS=-DDRL_TC_STAMPS python -c "import __graft_entry__ as g; g.build()"; python profiles/tc_stamps.py 4096"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DRL_TC_DEBUG"] = "1"
import deep_rl_b200 as drl  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = drl.PPOConfig(num_envs=N, num_steps=128, total_timesteps=N * 128 * 8)
tr = drl.PPOTrainer(cfg)
if os.environ.get("DRL_TC_PROBE"):   # the issuer waits for dh1 right after committing it and stamps the completion
    tr.workspace[-8:] = torch.tensor(list((77).to_bytes(8, "little")), dtype=torch.uint8, device=tr.workspace.device)
for _ in range(3):
    tr.update()
torch.cuda.synchronize()
ws = tr.workspace.cpu().numpy()
dbg = ws[-4096:].view(np.int64)
comp, mma = dbg[:256].reshape(16, 16), dbg[256:].reshape(16, 16)
t0 = comp[1, 0]
names_c = ["X start", "fwd ready", "heads done", "pair synced", "dz2 ready", "w1(k-1) ok", "BWD handed", "Z start(Y done)", "bwd ready", "W1 handed", "Z: before dh1 wait", "Z: dh1 wait passed"]
names_m = ["BWD sync", "bwd issued", "FWD sync", "fwd issued", "W1 sync", "w1 issued", "dh1 issued + committed"]
for k in range(1, 6):
    print(f"--- tile {k} (cycles relative to X start of tile 1)")
    ev = [(int(comp[k, i] - t0), "C " + names_c[i]) for i in range(12)] + [(int(mma[k, i] - t0), "M " + names_m[i]) for i in range(7) if mma[k, i] != 0]
    for t, n in sorted(ev):
        print(f"{t:8d}  {n}")

ks = [int(mma[n >> 3, 8 + (n & 7)]) for n in range(11)]
kn = ["kernel entry", "first record requested", "weights landed", "main loop start", "main loop end", "partials stored",
      "grid barrier 0 passed", "fold done", "norm share published", "grid barrier 2 passed", "Adam done"]
print("--- whole kernel, CTA 0 thread 0 (cycles from kernel entry)")
for n in range(11):
    print(f"{ks[n] - ks[0]:8d}  {kn[n]}")
