"""REINFORCE path (deep_rl/reinforce.py): oracle vs the unmodified script's own episodes (CPU), kernels vs the
script's episodes and vs the oracle (GPU)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import clib, reinforce_oracle as ro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = 500


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_reinforce.npz"))


def _planes(rew):
    """[T+1][1] reward / done planes with the one-slot shift and the after-episode padding of drl_reinforce_episodes."""
    L = len(rew)
    r = np.zeros((T + 1, 1), np.float32)
    d = np.ones((T + 1, 1), np.float32)
    r[1:L + 1, 0] = rew
    d[1:L, 0] = 0.0                      # done[L] = 1: the episode's last step
    d[0, 0] = 0.0
    return r, d


# ------------------------------------------------------------------------------------------------
# CPU: oracle against the unmodified script
# ------------------------------------------------------------------------------------------------
def test_oracle_matches_reference_episodes(ref):
    gamma = float(ref["gamma"])
    for i in range(4):
        g = lambda k: ref[f"e{i}_{k}"]
        L = len(g("act"))
        np.testing.assert_array_equal(ro.reward_to_go(g("rew"), gamma).numpy(), g("returns"))              # the script's accumulation, bit for bit
        loss, grad, b_ret, logp = ro.episode_loss_and_grad(g("params_before"), g("obs")[:L], g("act"), g("mask"), g("rew"), gamma)
        np.testing.assert_allclose(b_ret, g("b_returns"), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(logp, g("b_log_probs"), rtol=1e-6, atol=1e-7)
        assert abs(loss - float(g("policy_loss"))) <= 1e-5 * max(1.0, abs(loss))
        np.testing.assert_allclose(grad, g("grad"), rtol=1e-4, atol=1e-6)
        m = np.zeros_like(grad) if i == 0 else m                                                            # noqa: F821
        v = np.zeros_like(grad) if i == 0 else v                                                            # noqa: F821
        p, m, v = ro.adam(g("params_before"), g("grad"), m, v, i + 1)
        np.testing.assert_allclose(p, g("params_after"), rtol=0, atol=1e-7)


def test_reverse_scan_equals_the_scripts_reward_to_go(ref):
    """reinforce.py:67 == the GAE scan with lambda = 1 and V = 0, to fp32 rounding."""
    gamma = float(ref["gamma"])
    for i in range(4):
        rew = ref[f"e{i}_rew"]
        r, d = _planes(rew)
        adv, _ = clib.gae(r, d, np.zeros_like(r), gamma, 1.0)
        np.testing.assert_allclose(adv[:len(rew), 0], ref[f"e{i}_returns"], rtol=2e-6)
        assert np.all(adv[len(rew):, 0] == 0.0)
    rew = np.ones(500, np.float32)                                                                          # the longest possible episode
    r, d = _planes(rew)
    adv, _ = clib.gae(r, d, np.zeros_like(r), 0.99, 1.0)
    np.testing.assert_allclose(adv[:500, 0], ro.reward_to_go(rew, 0.99).numpy(), rtol=1e-5)


def test_mask_bit_packing_roundtrip():
    rng = np.random.default_rng(0)
    keep = rng.uniform(size=(7, 3, 128)) < 0.4
    assert np.array_equal(ro.unpack_mask_bits(ro.pack_mask_bits(keep)), keep)


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_kernels_teacher_forced_on_reference_episodes(ref):
    """The script's own episodes (observations, actions, dropout masks, rewards) through drl_gae(lambda=1), drl_reinforce_grad and
    drl_adam_step: returns, loss, gradient and parameters after the step against what the script computed."""
    import deep_rl_b200 as drl
    from deep_rl_b200 import _lib as L
    from deep_rl_b200.reinforce import ReinforceConfig, ReinforceTrainer
    tr = ReinforceTrainer(ReinforceConfig(num_envs=1))
    np.testing.assert_array_equal(tr.params.cpu().numpy(), ref["e0_params_before"])           # same init draws as the script (seed 1)
    gamma = float(ref["gamma"])
    for i in range(4):
        g = lambda k: ref[f"e{i}_{k}"]
        n = len(g("act"))
        tr.params.copy_(torch.tensor(g("params_before")))
        r, d = _planes(g("rew"))
        tr.rewards.copy_(torch.tensor(r)); tr.dones.copy_(torch.tensor(d.astype(np.uint8)))
        tr.observations.zero_(); tr.observations[:n + 1, 0] = torch.tensor(g("obs"))
        tr.actions.zero_(); tr.actions[:n, 0] = torch.tensor(g("act").astype(np.uint8))
        tr.ep_len.fill_(n)
        masks = torch.zeros((T, 1, 4), dtype=torch.int32, device=tr.device)
        masks[:n, 0] = torch.tensor(ro.pack_mask_bits(g("mask")).view(np.int32))
        tr.compute_returns()
        np.testing.assert_allclose(tr.returns[:n, 0].cpu().numpy(), g("returns"), rtol=2e-6)
        tr.adam_step = i
        tr.optimize(teacher_masks=masks)
        torch.cuda.synchronize()
        assert abs(float(tr.loss.item()) - float(g("policy_loss"))) <= 2e-5 * max(1.0, abs(float(g("policy_loss"))))
        np.testing.assert_allclose(tr.grad.cpu().numpy(), g("grad"), rtol=2e-4, atol=2e-6)
        if i == 0:                       # Adam from zero moments: the script's first step
            np.testing.assert_allclose(tr.params.cpu().numpy(), g("params_after"), rtol=0, atol=2e-7)
        tr.exp_avg.zero_(); tr.exp_avg_sq.zero_()


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 37, 256])
def test_episode_kernel_vs_oracle(N):
    """drl_reinforce_episodes: env dynamics of every episode against the oracle env driven by the kernel's actions; planes padded
    after the episode end; the kernel's Philox dropout masks (debug plane) give, through the ORACLE policy and the oracle sampler,
    exactly the kernel's actions; the gradient with regenerated masks equals the gradient with the same masks passed in."""
    from deep_rl_b200.reinforce import ReinforceConfig, ReinforceTrainer
    seed = 5
    tr = ReinforceTrainer(ReinforceConfig(num_envs=N, seed=seed), debug_masks=True)
    for it in range(2):
        step0 = tr.step0
        p0 = tr.params.cpu().numpy().copy()
        tr.episodes()
        tr.compute_returns()
        torch.cuda.synchronize()
        obs, act = tr.observations.cpu().numpy(), tr.actions.cpu().numpy()
        rew, done, ln = tr.rewards.cpu().numpy(), tr.dones.cpu().numpy(), tr.ep_len.cpu().numpy()
        keep = ro.unpack_mask_bits(tr.mask_bits.cpu().numpy().view(np.uint32))
        ret = tr.returns.cpu().numpy()
        assert 0.3 < keep[0].mean() < 0.5                                                    # Bernoulli(0.4)
        for n in range(min(N, 24)):
            L = int(ln[n])
            assert 8 <= L <= T
            st = np.zeros(4)
            clib.lib().drl_or_reset_state(0, seed, n, step0 - 1, st.ctypes.data_as(C.POINTER(C.c_double)))
            np.testing.assert_allclose(obs[0, n], st.astype(np.float32), atol=1e-7)
            for t in range(L):
                pr = ro.probs(torch.tensor(p0), torch.tensor(obs[t, n]), torch.tensor(keep[t, n]))
                logits = torch.log(pr).numpy()[None]                                           # the sampler is shift-invariant
                u = clib.action_uniform(seed, n, step0 + t)
                cdf0 = float(pr[0])
                if abs(u - cdf0) > 1e-5:                                                      # away from the CDF edge: the action is forced
                    assert act[t, n] == (0 if u < cdf0 else 1), (n, t, u, cdf0)
                st, term = clib.cartpole_step(st, int(act[t, n]))
                assert rew[t + 1, n] == 1.0 and done[t + 1, n] == (1 if (term or t + 1 >= 500) else 0)
                if t + 1 < L:
                    np.testing.assert_allclose(obs[t + 1, n], st.astype(np.float32), atol=1e-6)
            assert done[L, n] == 1 and np.all(rew[L + 1:, n] == 0) and np.all(done[L + 1:, n] == 1) and np.all(ret[L:, n] == 0)
            np.testing.assert_allclose(ret[:L, n], ro.reward_to_go(np.ones(L, np.float32), 0.99).numpy(), rtol=1e-5)
        # gradient: regenerated masks == the same masks passed in, and == torch autograd for a few episodes
        tr.optimize()
        torch.cuda.synchronize()
        g_regen = tr.grad.cpu().numpy().copy()
        part = tr._grad_part.cpu().numpy().copy()
        tr.params.copy_(torch.tensor(p0)); tr.adam_step -= 1; tr.exp_avg.zero_(); tr.exp_avg_sq.zero_()
        tr.optimize(teacher_masks=tr.mask_bits)
        torch.cuda.synchronize()
        assert np.array_equal(tr.grad.cpu().numpy(), g_regen)
        for n in range(min(N, 4)):
            L = int(ln[n])
            _, wg, _, _ = ro.episode_loss_and_grad(p0, obs[:L, n], act[:L, n], keep[:L, n], np.ones(L, np.float32), 0.99)
            np.testing.assert_allclose(part[n, :ro.P], wg, rtol=5e-4, atol=5e-6)
        tr.step0 += tr.T
        cnt, sum_ret, sum_len, entries = tr.env.log.drain()
        assert cnt == N and sum_len == float(ln.sum()) and sum_ret == sum_len


@pytest.mark.gpu
def test_reinforce_learns_cartpole(ref, capsys):
    """The reference goes from ~24 to ~160 in 100 episodes (golden).  16 envs x 100 iterations here: the mean return of the last ten
    iterations must exceed 100; with one env the printed lines have the script's format."""
    from deep_rl_b200.reinforce import ReinforceConfig, ReinforceTrainer, train
    want = ref["episodes"][:, 1]
    assert want[:10].mean() < 40 and want[-10:].mean() > 100
    tr = ReinforceTrainer(ReinforceConfig(num_envs=16, seed=1))
    rets = []
    for _ in range(100):
        tr.iteration()
        rets.append(tr.metrics()["mean_return"])
    assert np.mean(rets[:5]) < 40 and np.mean(rets[-10:]) > 100, (rets[:5], rets[-10:])
    train(ReinforceConfig(num_envs=1, num_episodes=5))
    out = capsys.readouterr().out.strip().splitlines()
    import re
    assert len(out) == 5 and all(re.fullmatch(r"global_step=\d+, episodic_return=\d+\.\d\d", ln) for ln in out)
