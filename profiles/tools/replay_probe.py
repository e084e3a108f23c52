"""Replay-buffer kernels on a 4 Mi-transition buffer: uniform draw + gather and prioritized draw + gather of 65,536-sample batches."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import deep_rl_b200 as drl
size, batch = 1 << 22, 1 << 16
for pri in (False, True):
    rb = drl.ReplayBuffer(size, 4, seed=1, prioritized=pri)
    rb.observations.normal_(); rb.rewards.normal_(); rb.size = size
    if pri:
        rb.priorities.uniform_(1e-3, 1.0)
    for _ in range(3):
        rb.sample(batch)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        rb.sample(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"prioritized={pri}: sample+gather of {batch} transitions from {size}: {ms * 1e3:.1f} us per batch "
          f"({batch / ms / 1e6:.2f} G transitions/s; gather algorithmic {41 * batch / ms / 1e6:.0f} GB/s)")
