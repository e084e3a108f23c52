"""Stand-alone drl_env_step (state in HBM) on 4 Mi CartPole envs with random actions: achieved HBM GB/s (94 B per env-step)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import deep_rl_b200 as drl
from deep_rl_b200 import _lib
for env_id, nbytes in (("CartPole-v1", 94), ("Acrobot-v1", 102)):
    n = 1 << 22
    env = drl.make(env_id, num_envs=n, seed=3, log_capacity=1 << 22)
    env.reset()
    act = torch.randint(0, env.num_actions, (n,), dtype=torch.int32, device=env.device)
    call = lambda k: _lib.check(env.L.drl_env_step(C.byref(env.struct), k, act.data_ptr(), env._obs.data_ptr(), env._rew.data_ptr(),
                                                   env._done.data_ptr(), C.byref(env.log.struct), _lib.stream_ptr()))
    for k in range(3):
        call(k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(20):
        call(3 + k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{env_id}: env_step_kernel {ms * 1e3:.1f} us for {n} envs, {nbytes * n / ms / 1e6:.0f} GB/s of the {nbytes} B/env-step ({nbytes * n / ms / 1e6 / 6552:.3f} of 6552 GB/s)")
    del env
