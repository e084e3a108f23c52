"""Per-step cycle timeline of rollout_tc_kernel (CTA 0).  Needs a library built with -DDRL_ROLLOUT_STAMPS:
   DRL_EXTRA_NVCC_FLAGS=-DDRL_ROLLOUT_STAMPS python -c "import __graft_entry__ as g; g.build()"; python profiles/tools/ro_stamps.py 4096"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import deep_rl_b200 as drl
from deep_rl_b200 import _lib as L
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tr = drl.PPOTrainer(drl.PPOConfig(num_envs=N, num_steps=128, total_timesteps=N * 128 * 8))
for _ in range(3):
    tr.update()
torch.cuda.synchronize()
out = np.zeros(1024, dtype=np.int64)
lib = L.lib()
lib.drl_debug_rollout_stamps.argtypes = [C.c_void_p]
assert lib.drl_debug_rollout_stamps(out.ctypes.data) == 0
names = ["S0 start", "S0 done", "group barrier 1 passed", "S1 done (tile stored)", "fwd handed", "fwd ready", "S2 done", "group barrier 2 passed",
         "S3: logits summed", "S3: philox done", "S3: sampled", "S3: env stepped"]
own, oth = out[:512].reshape(32, 16), out[512:].reshape(32, 16)
t0 = own[0, 0]
for k in range(3):
    print(f"--- step {8 + k}")
    ev = [(int(own[k, i] - t0), "owner  " + names[i]) for i in range(12)] + [(int(oth[k, i] - t0), "warp 5 " + names[i]) for i in range(8)]
    for t, n in sorted(ev):
        print(f"{t:8d}  {n}")
