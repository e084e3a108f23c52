"""Determinism stress: the same two updates repeated many times must give bit-identical parameters, with and without the
side-stream overlap.  Usage: python profiles/tools/stress_overlap.py [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import deep_rl_b200 as drl  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for env_id, N, T, prec in [("Acrobot-v1", 24, 64, "fp32"), ("CartPole-v1", 64, 32, "fp32"), ("CartPole-v1", 64, 32, "bf16"),
                           ("CartPole-v1", 4096, 128, "bf16")]:
    ref = {}
    bad = 0
    for overlap in (False, True):
        for rep in range(reps):
            cfg = drl.PPOConfig(env_id=env_id, num_envs=N, num_steps=T, seed=2, total_timesteps=N * T * 10, update_precision=prec,
                                overlap_streams=overlap)
            tr = drl.PPOTrainer(cfg)
            for _ in range(3):
                tr.update(10)
            torch.cuda.synchronize()
            p = tr.agent.flat_params.cpu().numpy().copy()
            lt = tr.loss_terms.cpu().numpy().copy()
            if "p" not in ref:
                ref["p"], ref["lt"] = p, lt
            elif not (np.array_equal(p, ref["p"]) and np.array_equal(lt, ref["lt"])):
                bad += 1
                print(f"  MISMATCH {env_id} N={N} {prec} overlap={overlap} rep={rep}: max |dp| = {np.abs(p - ref['p']).max():.3e}, "
                      f"loss-term rows differing: {np.nonzero((lt != ref['lt']).any(axis=1))[0][:8]}")
    print(f"{env_id} N={N} T={T} {prec}: {bad} mismatches in {2 * reps} runs")
