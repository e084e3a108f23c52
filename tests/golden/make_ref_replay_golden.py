"""Generate tests/golden/ref_replay.npz from the UNMODIFIED reference scripts deep_rl/dqn.py and deep_rl/per.py.

Run once, in the authoring container (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_ref_replay_golden.py

How: runpy.run_path on each script with oracle/gym_shim as `gym`.  The scripts are not edited; their replay-buffer path is
observed by wrapping the library calls that bracket it -- `np.random.randint` (dqn.py:116) / `torch.multinomial` (per.py:129)
deliver the batch indices, and `torch.optim.Adam.step` (dqn.py:133 / per.py:158) is the moment at which the script's globals hold
the gathered batch (`b_observations`, `b_actions`, `b_next_observations`, `b_rewards`, `b_terminated`, per.py also
`b_probabilities`, `weights`, the updated `priorities` and `max_priority`).  After a few training batches the run is stopped by an
exception raised from the wrapper (the remaining ~90,000 environment steps add nothing).

per.py hard-codes env_id = "LunarLander-v2" (Box2D, not installable here).  Its replay path -- priorities, multinomial draw,
probabilities, importance weights, priority update -- does not depend on the environment, so for that script `gym.make` is
wrapped to hand out CartPole-v1 instead; the script file itself is executed unmodified.
"""
import contextlib
import io
import os
import runpy
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "gym_shim"))
sys.path.insert(0, ROOT)
BATCHES = 3


class _Enough(Exception):
    pass


def _script_globals(marker):
    f = sys._getframe(1)
    while f is not None:
        g = f.f_globals
        if marker in g and g.get("__name__") == "__main__":
            return g
        f = f.f_back
    raise RuntimeError("reference frame not found")


def _np(t):
    return t.detach().cpu().numpy().copy()


def run_dqn():
    rec = {"idx": [], "batches": []}
    orig_randint, orig_step = np.random.randint, torch.optim.Adam.step

    def randint_spy(*a, **k):
        out = orig_randint(*a, **k)
        if "size" in k and k["size"] == 128:          # dqn.py:116 (env.action_space.sample() goes through the shim's own RandomState)
            rec["idx"].append(np.array(out, dtype=np.int64))
        return out

    def step_spy(self, *a, **k):
        g = _script_globals("batch_inds")
        n = int(g["global_step"])
        rec["batches"].append({
            "global_step": n, "batch_inds": np.array(g["batch_inds"], dtype=np.int64),
            "observations": _np(g["observations"][: n + 1]), "actions": _np(g["actions"][: n + 1]),
            "rewards": _np(g["rewards"][: n + 1]), "terminated": _np(g["terminated"][: n + 1]),
            "b_observations": _np(g["b_observations"]), "b_actions": _np(g["b_actions"]),
            "b_next_observations": _np(g["b_next_observations"]), "b_rewards": _np(g["b_rewards"]),
            "b_terminated": _np(g["b_terminated"])})
        out = orig_step(self, *a, **k)
        if len(rec["batches"]) >= BATCHES:
            raise _Enough()
        return out

    np.random.randint, torch.optim.Adam.step = randint_spy, step_spy
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            runpy.run_path("/root/reference/deep_rl/dqn.py", run_name="__main__")
    except _Enough:
        pass
    finally:
        np.random.randint, torch.optim.Adam.step = orig_randint, orig_step
    for i, b in enumerate(rec["batches"]):
        assert np.array_equal(b["batch_inds"], rec["idx"][i])
    return rec["batches"]


def run_per():
    rec = {"pre": [], "batches": []}
    orig_multinomial, orig_step = torch.multinomial, torch.optim.Adam.step

    def multinomial_spy(inp, num, *a, **k):
        out = orig_multinomial(inp, num, *a, **k)
        rec["pre"].append({"priorities_before": _np(inp), "batch_inds": _np(out)})
        return out

    def step_spy(self, *a, **k):
        g = _script_globals("batch_inds")
        n = int(g["global_step"])
        pre = rec["pre"][len(rec["batches"])]
        assert np.array_equal(pre["batch_inds"], _np(g["batch_inds"]))
        rec["batches"].append({
            "global_step": n, "batch_inds": pre["batch_inds"], "priorities_before": pre["priorities_before"][: n + 1],
            "alpha": float(g["alpha"]), "beta": float(g["beta"]),
            "b_probabilities": _np(g["b_probabilities"]), "td_errors": _np(g["td_errors"]), "weights": _np(g["weights"]),
            "priorities_after": _np(g["priorities"][: n + 1]), "max_priority": float(g["max_priority"]),
            "b_next_observations": _np(g["b_next_observations"]), "b_rewards": _np(g["b_rewards"]),
            "observations": _np(g["observations"][: n + 2]), "rewards": _np(g["rewards"][: n + 2])})
        out = orig_step(self, *a, **k)
        if len(rec["batches"]) >= BATCHES:
            raise _Enough()
        return out

    import gym
    orig_make = gym.make
    gym.make = lambda env_id: orig_make("CartPole-v1")          # see the module docstring
    torch.multinomial, torch.optim.Adam.step = multinomial_spy, step_spy
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            runpy.run_path("/root/reference/deep_rl/per.py", run_name="__main__")
    except _Enough:
        pass
    finally:
        torch.multinomial, torch.optim.Adam.step = orig_multinomial, orig_step
        gym.make = orig_make
    return rec["batches"]


def main():
    out = {}
    for name, batches in (("dqn", run_dqn()), ("per", run_per())):
        for i, b in enumerate(batches):
            for k, v in b.items():
                out[f"{name}{i}_{k}"] = np.asarray(v)
        print(name, "batches:", [(int(b["global_step"]), b["batch_inds"][:4].tolist()) for b in batches])
    path = os.path.join(ROOT, "tests", "golden", "ref_replay.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
