"""GPU: size-independent properties at BASELINE.json's full sizes (C2: 4096 envs x 128 steps, C3: 65,536 envs x 128 steps
per GPU) -- the oracle is too slow to replay these end to end, so the checks are invariants of the domain plus oracle
replays of a random subset of environments."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import clib  # noqa: E402

THETA = 12 * 2 * np.pi / 360


@pytest.mark.parametrize("N,precision", [(4096, "fp32"), (65_536, "bf16")])
def test_rollout_invariants_and_subset_replay(N, precision):
    import deep_rl_b200 as drl
    T, seed = 128, 1
    cfg = drl.PPOConfig(num_envs=N, num_steps=T, seed=seed, rollout_precision=precision, total_timesteps=N * T * 8)
    tr = drl.PPOTrainer(cfg)
    for _ in range(3):            # third rollout: envs are in every phase of their episodes
        tr.rollout()
    torch.cuda.synchronize()
    obs = tr.observations[:, :, :4]
    act, logp, rew, done = tr.actions, tr.log_probs, tr.rewards, tr.dones
    assert bool((rew[1:] == 1.0).all())                                  # CartPole reward
    assert int(act[:T].max()) <= 1 and bool((logp[:T] <= 0).all()) and bool(torch.isfinite(tr.values).all())
    d = done[1:].bool()
    nxt = obs[1:]
    inside = (nxt[..., 0].abs() <= 2.4) & (nxt[..., 2].abs() <= THETA)
    assert bool(inside.all())                                            # stored observations are never terminal ones
    reset_like = (nxt.abs() <= 0.05).all(-1)
    assert bool(reset_like[d].all())                                     # after done the stored obs is a fresh reset draw
    cnt, sum_ret, sum_len, _ = tr.env.log.drain(with_entries=False)
    assert sum_ret == sum_len and cnt >= int(d.sum())                    # return == length for CartPole (3 rollouts of episodes)
    np.testing.assert_allclose(tr.env.observe().cpu().numpy(), tr.observations[T, :, :4].cpu().numpy(), atol=0)
    # oracle replay of 64 random envs over the last rollout, driven by the kernel's own actions
    rng = np.random.default_rng(0)
    cols = np.sort(rng.choice(N, size=64, replace=False))
    o_np, a_np, r_np, d_np = (x.cpu().numpy() for x in (obs, act, rew, done))
    for n in cols[:16]:
        st = np.zeros(4)
        # recover the float64 state at t=0 is impossible from float32 obs, so replay from the first reset inside the window
        starts = np.nonzero(d_np[1:, n])[0]
        if len(starts) == 0:
            continue
        t0 = int(starts[0]) + 1                                          # obs[t0] is a reset draw: regenerate it exactly
        ora = clib.OracleVecEnv("CartPole-v1", 1, seed=seed, env_gid0=int(n))
        clib.lib().drl_or_reset_state(0, seed, int(n), 2 * T + t0 - 1, ora.state.ctypes.data_as(C.POINTER(C.c_double)))
        np.testing.assert_allclose(ora.observe()[0], o_np[t0, n], atol=1e-7)
        ora.step_count = 2 * T + t0
        for t in range(t0, T):
            o, r, dd, _ = ora.step(a_np[t, n:n + 1])
            np.testing.assert_allclose(o[0], o_np[t + 1, n], atol=1e-6, err_msg=f"env {n} t={t}")
            assert dd[0] == d_np[t + 1, n]


def test_gae_full_size_subset_vs_oracle():
    import deep_rl_b200 as drl
    N, T = 65_536, 128
    cfg = drl.PPOConfig(num_envs=N, num_steps=T, seed=2, total_timesteps=N * T * 8)
    tr = drl.PPOTrainer(cfg)
    tr.rollout()
    tr.compute_gae()
    torch.cuda.synchronize()
    cols = np.sort(np.random.default_rng(1).choice(N, size=512, replace=False))
    rew, done, val = (getattr(tr, k)[:, cols].cpu().numpy() for k in ("rewards", "dones", "values"))
    adv, ret = clib.gae(rew, done.astype(np.float32), val, cfg.gamma, cfg.gae_lambda)
    assert np.array_equal(tr.advantages[:, cols].cpu().numpy(), adv)
    assert np.array_equal(tr.returns[:, cols].cpu().numpy(), ret)
    # records carry exactly the plane values
    rec = tr.records.view(T, N, -1)[:, cols].cpu().numpy()
    assert np.array_equal(rec[..., 5], adv[:T]) and np.array_equal(rec[..., 6], val[:T])
    assert np.array_equal(rec[..., 7].view(np.int32), tr.actions[:T, cols].cpu().numpy().astype(np.int32))


def test_permutation_bijection_c3_size():
    from deep_rl_b200 import _lib as L
    B = 65_536 * 128
    idx = torch.empty(B, dtype=torch.int32, device="cuda:0")
    L.check(L.lib().drl_permutation(idx.data_ptr(), B, 1, 5, 0, L.stream_ptr()))
    s, _ = torch.sort(idx.to(torch.int64))
    assert torch.equal(s, torch.arange(B, device="cuda:0"))
    head = idx[:2048].cpu().numpy().view(np.uint32)
    want = np.array([clib.lib().drl_or_perm_index(i, B, 1, 5, 0) for i in range(2048)], dtype=np.uint32)
    assert np.array_equal(head, want)


def test_update_full_size_deterministic_and_precisions_agree():
    """One whole C2 update twice from the same state: bit-identical; bf16 vs fp32 update within bf16 tolerance."""
    import deep_rl_b200 as drl
    res = {}
    for prec in ("bf16", "bf16", "fp32"):
        cfg = drl.PPOConfig(num_envs=4096, num_steps=128, seed=7, total_timesteps=4096 * 128 * 8, update_precision=prec,
                            rollout_precision="fp32")
        tr = drl.PPOTrainer(cfg)
        p0 = tr.agent.flat_params.clone()
        tr.update()
        torch.cuda.synchronize()
        res.setdefault(prec, []).append(((tr.agent.flat_params - p0).cpu().numpy(), tr.loss_terms.cpu().numpy().copy()))
    (d1, l1), (d2, l2) = res["bf16"]
    assert np.array_equal(d1, d2) and np.array_equal(l1, l2)
    d32, l32 = res["fp32"][0]
    np.testing.assert_allclose(l1[:, :4], l32[:, :4], rtol=3e-2, atol=1e-2)
    assert np.linalg.norm(d1 - d32) / np.linalg.norm(d32) < 0.2
