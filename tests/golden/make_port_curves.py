"""Generate tests/golden/port_curves_10seeds.npz: learning curves of the reference's algorithm on the CPU for ten seeds.

Run once, in the authoring container:   python tests/golden/make_port_curves.py

The reference script hard-codes seed = 1 (deep_rl/ppo.py:83), so the other seeds come from oracle/ppo_port.py, the single-env
port that reproduces the UNMODIFIED script bit for bit at seed 1 (checked here against tests/golden/ref_ppo_seed1.npz: same
episode sequence, same returns).  Per seed: the return of every finished episode of the 20,000-step run (156 updates).
The GPU build uses different RNG streams by contract, so its curves can only agree with these in
DISTRIBUTION: tests/test_gpu_rollout.py::test_learning_curve_distribution_matches_the_port compares the last-20-episode
means of ten seeds of each.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ppo_port as pp  # noqa: E402

SEEDS = list(range(1, 11))


def main():
    golden = np.load(os.path.join(ROOT, "tests", "golden", "ref_ppo_seed1.npz"))
    out = {"seeds": np.array(SEEDS)}
    for s in SEEDS:
        tr = pp.run(pp.PortConfig(seed=s))
        eps = np.array(tr.episodes, dtype=np.float64)
        if s == 1:      # the port IS the reference at the reference's seed
            assert np.array_equal(eps, golden["episodes"]), "port diverged from the unmodified reference at seed 1"
        out[f"seed{s}_episodes"] = eps
        print(f"seed {s}: {len(eps)} episodes, first-20 mean {eps[:20, 1].mean():.1f}, last-20 mean {eps[-20:, 1].mean():.1f}")
    out["first20"] = np.array([out[f"seed{s}_episodes"][:20, 1].mean() for s in SEEDS])
    out["last20"] = np.array([out[f"seed{s}_episodes"][-20:, 1].mean() for s in SEEDS])
    path = os.path.join(ROOT, "tests", "golden", "port_curves_10seeds.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "last20:", np.round(out["last20"], 1), "mean", out["last20"].mean().round(1), "std", out["last20"].std(ddof=1).round(1))


if __name__ == "__main__":
    main()
