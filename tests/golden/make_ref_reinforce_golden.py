"""Generate tests/golden/ref_reinforce.npz from the UNMODIFIED reference script deep_rl/reinforce.py.

Run once, in the authoring container:    python tests/golden/make_ref_reinforce_golden.py

How: runpy.run_path with oracle/gym_shim as `gym`.  The script is not edited; it is observed through
  * a recording wrapper around the environment `gym.make` hands out (observations and actions of every step),
  * `torch.nn.functional.dropout` (the keep mask nn.Dropout drew from torch's CPU stream: output != 0),
  * `torch.optim.Adam.step` (at that moment the script's globals hold `returns`, `b_returns`, `b_log_probs`, `policy_loss`, the
    parameters carry their gradients; parameters before / after the step are read around the original call).
Stored: the first episodes in full, and the (global_step, episodic_return) of all 100 episodes.
"""
import contextlib
import io
import os
import re
import runpy
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "gym_shim"))
sys.path.insert(0, ROOT)
FULL = 4


def main():
    import gym
    rec = {"eps": [], "cur_obs": [], "cur_act": [], "cur_rew": [], "cur_mask": []}
    orig_make, orig_dropout, orig_step = gym.make, torch.nn.functional.dropout, torch.optim.Adam.step

    class Recorder(gym.Wrapper):
        def reset(self, **kw):
            o = self.env.reset(**kw)
            rec["cur_obs"], rec["cur_act"], rec["cur_rew"], rec["cur_mask"] = [np.array(o, dtype=np.float32)], [], [], []
            return o

        def step(self, action):
            o, r, d, info = self.env.step(action)
            rec["cur_act"].append(int(np.asarray(action).item()))
            rec["cur_rew"].append(float(r))
            rec["cur_obs"].append(np.array(o, dtype=np.float32))
            return o, r, d, info

    def make_spy(env_id):
        env = Recorder(orig_make(env_id))
        return env

    def dropout_spy(inp, p=0.5, training=True, inplace=False):
        out = orig_dropout(inp, p, training, inplace)
        rec["cur_mask"].append((out.detach() != 0).numpy().copy() | (inp.detach() == 0).numpy())
        return out

    def step_spy(self, *a, **k):
        f = sys._getframe(1)
        while f is not None and not ("b_returns" in f.f_globals and f.f_globals.get("__name__") == "__main__"):
            f = f.f_back
        g = f.f_globals
        params = list(g["agent"].parameters())
        flat = lambda grad: torch.cat([(q.grad if grad else q.detach()).reshape(-1) for q in params]).numpy().copy()
        entry = None
        if len(rec["eps"]) < FULL:
            L = int(g["step"])
            entry = {"obs": np.stack(rec["cur_obs"]), "act": np.array(rec["cur_act"]), "rew": np.array(rec["cur_rew"], np.float32),
                     "mask": np.stack(rec["cur_mask"]), "returns": g["returns"][:L].detach().numpy().copy(),
                     "b_returns": g["b_returns"].detach().numpy().copy(), "b_log_probs": g["b_log_probs"].detach().numpy().copy(),
                     "policy_loss": float(g["policy_loss"]), "grad": flat(True), "params_before": flat(False)}
        out = orig_step(self, *a, **k)
        if entry is not None:
            entry["params_after"] = flat(False)
            rec["eps"].append(entry)
        return out

    gym.make, torch.nn.functional.dropout, torch.optim.Adam.step = make_spy, dropout_spy, step_spy
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            g = runpy.run_path("/root/reference/deep_rl/reinforce.py", run_name="__main__")
    finally:
        gym.make, torch.nn.functional.dropout, torch.optim.Adam.step = orig_make, orig_dropout, orig_step
    eps = np.array([(int(m.group(1)), float(m.group(2))) for m in re.finditer(r"global_step=(\d+), episodic_return=([0-9.]+)", buf.getvalue())])
    out = {"episodes": eps, "gamma": np.array(g["gamma"]), "seed": np.array(g["seed"])}
    for i, e in enumerate(rec["eps"]):
        assert len(e["act"]) == len(e["mask"]) == len(e["returns"]) and len(e["obs"]) == len(e["act"]) + 1
        for k, v in e.items():
            out[f"e{i}_{k}"] = np.asarray(v)
    path = os.path.join(ROOT, "tests", "golden", "ref_reinforce.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(eps), "episodes, returns first 5", eps[:5, 1], "last 5", eps[-5:, 1],
          "lengths stored", [len(e["act"]) for e in rec["eps"]])


if __name__ == "__main__":
    main()
