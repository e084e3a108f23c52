"""Replay-buffer path: oracle vs the unmodified reference scripts' own batches (CPU), kernels vs oracle (GPU)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import clib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_replay.npz"))


# ------------------------------------------------------------------------------------------------
# CPU: the oracle restatement against what dqn.py / per.py gathered themselves
# ------------------------------------------------------------------------------------------------
def test_oracle_gather_matches_dqn_batches(ref):
    for i in range(3):
        g = lambda k: ref[f"dqn{i}_{k}"]
        b_obs, b_act, b_next, b_rew, b_term = clib.replay_gather(g("observations"), g("actions"), g("rewards"), g("terminated"), g("batch_inds"))
        assert np.array_equal(b_obs, g("b_observations")) and np.array_equal(b_next, g("b_next_observations"))
        assert np.array_equal(b_act, g("b_actions")) and np.array_equal(b_rew, g("b_rewards")) and np.array_equal(b_term, g("b_terminated"))
        assert g("batch_inds").max() < g("global_step")                  # np.random.randint(global_step, ...): never the open slot


def test_oracle_per_probabilities_and_update_match_reference(ref):
    for i in range(3):
        g = lambda k: ref[f"per{i}_{k}"]
        prob = clib.replay_probabilities(g("priorities_before"), g("batch_inds"), float(g("alpha")))
        np.testing.assert_allclose(prob, g("b_probabilities"), rtol=1e-6)
        w = (float(g("global_step")) * prob) ** -float(g("beta"))
        np.testing.assert_allclose(w / w.max(), g("weights"), rtol=1e-5)
        after = g("priorities_before").copy()
        after[g("batch_inds")] = np.abs(g("td_errors"))                  # numpy: last duplicate wins, like the reference's index_put on CPU
        assert np.array_equal(after, g("priorities_after"))
        assert float(g("max_priority")) == max(float(after.max()), 1e-2 if i == 0 else float(ref[f"per{i - 1}_max_priority"]))


def test_oracle_samplers_contract():
    idx = clib.replay_uniform(200_000, 10_000, seed=1, draw_ctr=0)
    assert idx.min() >= 0 and idx.max() < 10_000
    counts = np.bincount(idx, minlength=10_000)
    assert abs(counts.mean() - 20.0) < 1e-9 and counts.std() < 6.0           # uniform: Poisson(20) spread
    assert not np.array_equal(idx[:1000], clib.replay_uniform(1000, 10_000, seed=1, draw_ctr=1))
    rng = np.random.default_rng(0)
    pri = rng.uniform(size=5000).astype(np.float32) ** 3
    pri[777] = 50.0
    got = clib.replay_priority(pri, 400_000, seed=3, draw_ctr=5)
    freq = np.bincount(got, minlength=5000) / 400_000
    want = pri.astype(np.float64) / pri.astype(np.float64).sum()
    assert abs(freq[777] - want[777]) < 0.01 and np.abs(freq - want).max() < 0.01
    assert freq[pri == 0].sum() == 0 if (pri == 0).any() else True


def test_library_exports_replay_symbols():
    from deep_rl_b200 import _lib as L
    lib = L.lib()
    assert lib.drl_replay_scratch_bytes(10_001) >= 8 * 2 * 11
    assert lib.drl_replay_sample_uniform(0, 4, 10, 1, 0, 0) == -1
    assert lib.drl_replay_gather(0, 0, 0, 0, 0, 4, 4, 0, 0, 0, 0, 0, 0) == -1


# ------------------------------------------------------------------------------------------------
# GPU: kernels behind the C ABI vs the oracle / the reference's batches
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_replay_gather_kernel_vs_reference_batches(ref):
    import deep_rl_b200 as drl
    for i in range(3):
        g = lambda k: ref[f"dqn{i}_{k}"]
        n = int(g("global_step"))
        rb = drl.ReplayBuffer(n + 8, 4, seed=1)
        rb.observations[: n + 1, :4] = torch.tensor(g("observations"))
        rb.actions[: n + 1] = torch.tensor(g("actions").astype(np.int32))
        rb.rewards[: n + 1] = torch.tensor(g("rewards"))
        rb.terminated[: n + 1] = torch.tensor(g("terminated").astype(np.uint8))
        rb.size = n
        b = rb.gather(torch.tensor(g("batch_inds")))
        assert np.array_equal(b["observations"].cpu().numpy(), g("b_observations"))
        assert np.array_equal(b["next_observations"].cpu().numpy(), g("b_next_observations"))
        assert np.array_equal(b["actions"].cpu().numpy(), g("b_actions").astype(np.int32))
        assert np.array_equal(b["rewards"].cpu().numpy(), g("b_rewards"))
        assert np.array_equal(b["terminated"].cpu().numpy().astype(bool), g("b_terminated"))


@pytest.mark.gpu
@pytest.mark.parametrize("size,batch,obs_dim", [(10_000, 128, 4), (1, 16, 4), (1_000_003, 65_536, 6), (5000, 1, 2)])
def test_replay_samplers_and_gather_bit_exact_vs_oracle(size, batch, obs_dim):
    import deep_rl_b200 as drl
    rng = np.random.default_rng(size)
    rb = drl.ReplayBuffer(size, obs_dim, seed=7, prioritized=True, alpha=0.6)
    OP = rb.obs_stride
    obs = rng.normal(size=(size + 1, OP)).astype(np.float32)
    obs[:, obs_dim:] = 0
    act = rng.integers(0, 3, size=size + 1).astype(np.int32)
    rew = rng.normal(size=size + 1).astype(np.float32)
    term = (rng.uniform(size=size + 1) < 0.1).astype(np.uint8)
    pri = (rng.uniform(size=size + 1).astype(np.float32) ** 4) + 1e-3
    rb.observations.copy_(torch.tensor(obs)); rb.actions.copy_(torch.tensor(act)); rb.rewards.copy_(torch.tensor(rew))
    rb.terminated.copy_(torch.tensor(term)); rb.priorities.copy_(torch.tensor(pri)); rb.size = size
    for draw in range(2):
        b = rb.sample(batch)
        idx = b["batch_inds"].cpu().numpy().view(np.uint32)
        assert np.array_equal(idx, clib.replay_priority(pri[:size], batch, 7, draw))          # prioritized draw: bit-exact
        np.testing.assert_allclose(b["probabilities"].cpu().numpy(), clib.replay_probabilities(pri[:size], idx, 0.6), rtol=2e-5)
        w_obs, w_act, w_next, w_rew, w_term = clib.replay_gather(obs, act, rew, term, idx)
        assert np.array_equal(b["observations"].cpu().numpy(), w_obs[:, :obs_dim]) and np.array_equal(b["next_observations"].cpu().numpy(), w_next[:, :obs_dim])
        assert np.array_equal(b["actions"].cpu().numpy(), w_act) and np.array_equal(b["rewards"].cpu().numpy(), w_rew)
        assert np.array_equal(b["terminated"].cpu().numpy(), w_term)
    uni = drl.ReplayBuffer(size, obs_dim, seed=9)
    uni.size = size
    for draw in range(2):
        assert np.array_equal(uni.sample_indices(batch).cpu().numpy().view(np.uint32), clib.replay_uniform(batch, size, 9, draw))


@pytest.mark.gpu
def test_priority_update_vs_reference(ref):
    import deep_rl_b200 as drl
    prev_max = 1e-2
    for i in range(3):
        g = lambda k: ref[f"per{i}_{k}"]
        n = int(g("global_step"))
        rb = drl.ReplayBuffer(n + 8, 4, prioritized=True)
        rb.priorities[: n + 1] = torch.tensor(g("priorities_before"))
        rb.size = n
        rb.max_priority.fill_(prev_max)
        rb.update_priorities(torch.tensor(g("batch_inds")), torch.tensor(g("td_errors")))
        assert np.array_equal(rb.priorities[: n + 1].cpu().numpy(), g("priorities_after"))
        assert float(rb.max_priority.item()) == np.float32(g("max_priority"))
        prev_max = float(g("max_priority"))
    # duplicates: the last occurrence wins and only written values can raise the maximum
    rb = drl.ReplayBuffer(16, 4, prioritized=True)
    rb.size = 16
    rb.update_priorities(torch.tensor([3, 5, 3, 7, 3]), torch.tensor([9.0, -2.0, 0.5, 1.0, -0.25]))
    p = rb.priorities.cpu().numpy()
    assert p[3] == 0.25 and p[5] == 2.0 and p[7] == 1.0 and float(rb.max_priority.item()) == 2.0
