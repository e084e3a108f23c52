"""Learning curves of the B200 PPO path next to the reference's recorded curve (tests/golden/ref_ppo_seed1.npz).

    python examples/learning_curve.py            # reference shape (1 env x 128 steps, 20k timesteps), seeds 1..5
    python examples/learning_curve.py 4096 60    # 4096 envs, 60 updates
    python examples/learning_curve.py 4096 60 Acrobot-v1
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_rl_b200 as drl  # noqa: E402


def reference_shape():
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ref_ppo_seed1.npz"))
    eps = g["episodes"]
    print(f"reference (unmodified ppo.py, seed 1): {len(eps)} episodes, first-20 mean {eps[:20, 1].mean():.1f}, last-20 mean {eps[-20:, 1].mean():.1f}")
    for prec in ("fp32", "bf16"):
        for seed in (1, 2, 3, 4, 5):
            cfg = drl.PPOConfig(seed=seed, update_precision=prec)
            tr = drl.PPOTrainer(cfg)
            rets = []
            for _ in range(cfg.num_updates()):
                tr.update()
                rets += [e[2] for e in tr.metrics()["episode_log"]]
            print(f"b200 {prec} seed {seed}: {len(rets)} episodes, first-20 mean {np.mean(rets[:20]):.1f}, last-20 mean {np.mean(rets[-20:]):.1f}")


def many_envs(n, updates, env_id="CartPole-v1"):
    cfg = drl.PPOConfig(env_id=env_id, num_envs=n, total_timesteps=n * 128 * updates, seed=1)
    tr = drl.PPOTrainer(cfg)
    for u in range(updates):
        tr.update()
        m = tr.metrics(with_episode_log=False)
        if u % max(1, updates // 15) == 0 or u == updates - 1:
            print(f"update {u:4d} global_step {tr.global_step:>12d} episodes {m['episodes']:>7d} mean_return {m['mean_return']:7.2f} "
                  f"loss {m['loss']:.3f} entropy {m['entropy']:.3f} kl {m['approx_kl']:.4f} clipfrac {m['clipfrac']:.3f}")


if __name__ == "__main__":
    if len(sys.argv) > 2:
        many_envs(int(sys.argv[1]), int(sys.argv[2]), *(sys.argv[3:4]))
    else:
        reference_shape()
