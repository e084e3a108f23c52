"""CPU: pins the oracle against (a) outputs of the UNMODIFIED reference script captured in
tests/golden/ref_ppo_seed1.npz, (b) Random123's Philox known-answer vectors, (c) the classic-control
known-answer vectors of SURVEY.md App. C."""
import math

import numpy as np
import pytest
import torch

from oracle import clib, ppo_oracle as po, ppo_port as pp

GAE_UPDATES = (0, 1, 2, 50, 155)


def test_philox_known_answers():
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        assert clib.philox4x32_10(ctr, key).tolist() == want


def test_cartpole_known_answers():
    want = [
        (0, 0.1951219512195122, 0, -0.29268292682926828),
        (0.0039024390243902443, 0.3902439024390244, -0.0058536585365853658, -0.58536585365853655),
        (0.011707317073170733, 0.19520443149030264, -0.017560975609756099, -0.29453262531977781),
        (0.015611405702976786, 0.39057228683543699, -0.023451628116151656, -0.59270188416132941),
        (0.023422851439685526, 0.19578634341052323, -0.035305665799378244, -0.3074974581415999),
        (0.027338578307895989, 0.0011847918022427073, -0.041455614962210244, -0.026154747983545223),
    ]
    s = np.zeros(4)
    for a, w in zip([1, 1, 0, 1, 0, 0], want):
        s, term = clib.cartpole_step(s, a)
        np.testing.assert_allclose(s, w, rtol=0, atol=1e-15)
        assert not term
    s = np.array([0.01, -0.02, 0.03, 0.04])
    for a in [0, 1, 1]:
        s, _ = clib.cartpole_step(s, a)
    np.testing.assert_allclose(s, (0.0048718496976496205, 0.1736941202746195, 0.038823538610649126, -0.22139198924408143), atol=1e-15)
    s, n = np.zeros(4), 0
    while True:
        s, term = clib.cartpole_step(s, 1)
        n += 1
        if term:
            break
    assert n == 9 and abs(s[2] - (-0.21518604988500967)) < 1e-14


def test_acrobot_known_answers():
    want = [
        (-0.013262967177227795, 0.034287229347385442, -0.12866185280996106, 0.33450108998660194),
        (-0.048489653809372597, 0.12748779480535294, -0.21328487327674689, 0.57546306404962333),
        (-0.067237063386205026, 0.18553771278886602, 0.030714304370452555, -0.0064477093097422555),
        (-0.049859025656954617, 0.15938587589964667, 0.13865703759659068, -0.24781201811290368),
    ]
    s = np.zeros(4)
    for a, w in zip([2, 2, 0, 1], want):
        s, r, term = clib.acrobot_step(s, a)
        np.testing.assert_allclose(s, w, rtol=0, atol=1e-14)
        assert r == -1.0 and not term
    s = np.array([0.05, -0.03, 0.02, -0.07])
    for a in [0, 2]:
        s, _, _ = clib.acrobot_step(s, a)
    np.testing.assert_allclose(s, (0.053516999920591918, -0.084417236757849096, -0.147654816493417, 0.17725646848363069), atol=1e-14)


def test_time_limit_and_episode_statistics():
    # an env that never terminates physically within 500 steps is cut at 500 (done=1) and auto-resets
    env = clib.OracleVecEnv("Acrobot-v1", 2, seed=3)
    env.reset()
    for t in range(500):
        obs, rew, done, info = env.step(np.ones(2, np.int32))   # torque 0: hangs near the bottom
        if t < 499:
            assert not done.any()
    assert done.all() and (info["final_length"] == 500).all() and (info["final_return"] == -500.0).all()
    assert (env.elapsed == 0).all() and (env.ep_len == 0).all()
    # CartPole: reward 1 every step incl. the terminating one; episode return == length
    env = clib.OracleVecEnv("CartPole-v1", 16, seed=5)
    env.reset()
    seen = 0
    for t in range(200):
        _, rew, done, info = env.step(np.ones(16, np.int32))
        assert (rew == 1.0).all()
        for i in np.nonzero(done)[0]:
            assert info["final_return"][i] == info["final_length"][i]
            seen += 1
    assert seen > 16


@pytest.mark.parametrize("upd", GAE_UPDATES)
def test_gae_bit_exact_vs_reference(golden, upd):
    adv, ret = clib.gae(golden[f"u{upd}_rewards"][:, None], golden[f"u{upd}_dones"][:, None],
                        golden[f"u{upd}_values"][:, None], 0.99, 0.95)
    assert np.array_equal(adv[:, 0], golden[f"u{upd}_advantages"])
    assert np.array_equal(ret[:, 0], golden[f"u{upd}_returns"])


def test_init_bit_exact_vs_reference(golden):
    assert np.array_equal(po.init_params(4, 64, 2, 1).numpy(), golden["init_params"])


def test_loss_grad_adam_vs_reference(golden):
    g = golden
    params = g["init_params"]
    m, v = np.zeros_like(params), np.zeros_like(params)
    for i in range(16):
        sel = g["perms"][i // 4][(i % 4) * 32:(i % 4 + 1) * 32]
        terms, grad = po.minibatch_loss_and_grad(params, g["u0_observations"][sel], g["u0_actions"][sel],
                                                 g["u0_log_probs"][sel], g["u0_advantages"][sel], g["u0_returns"][sel],
                                                 g["u0_values"][sel], 4, 64, 2)
        np.testing.assert_allclose(terms, g["mb_terms"][i], rtol=0, atol=5e-7)
        np.testing.assert_allclose(grad, g["mb_grad_pre"][i], rtol=0, atol=1e-7)
        p2, m, v, norm = po.clip_adam(params, g["mb_grad_pre"][i], m, v, i + 1, float(g["mb_lr"][i]))
        assert abs(norm - g["mb_norm"][i]) < 1e-5
        np.testing.assert_allclose(p2, g["mb_params_after"][i], rtol=0, atol=2e-7)
        params = g["mb_params_after"][i]


def test_port_bit_identical_to_reference(golden):
    snaps = {}

    def cb(u, d):
        snaps[u] = {k: v.detach().numpy().copy() for k, v in d.items() if k != "agent"}
        snaps[u]["params"] = torch.cat([p.detach().reshape(-1) for p in d["agent"].parameters()]).numpy().copy()

    tr = pp.run(max_updates=3, on_update=cb)
    names = dict(obs="observations", val="values", act="actions", logp="log_probs", rew="rewards", done="dones",
                 adv="advantages", ret="returns")
    for u in (0, 1, 2):
        for k, gk in names.items():
            np.testing.assert_allclose(snaps[u][k], golden[f"u{u}_{gk}"], rtol=0, atol=1e-6, err_msg=f"update {u} {k}")
    np.testing.assert_allclose(snaps[0]["params"], golden["mb_params_after"][15], rtol=0, atol=1e-6)
    eps = golden["episodes"]
    k = len(tr.episodes)
    assert k > 5 and [e[0] for e in tr.episodes] == eps[:k, 0].astype(int).tolist()


def test_reference_learning_curve_recorded(golden):
    eps = golden["episodes"]
    assert len(eps) == 179 and eps[:20, 1].mean() < 40 and eps[-20:, 1].mean() > 150


def test_exp_det_accuracy():
    for x in np.linspace(-87, 0, 2001):
        want = math.exp(float(np.float32(x)))
        got = clib.exp_det(float(x))
        assert abs(got - want) <= 2.5e-7 * want + 1e-45
    assert clib.exp_det(-100.0) == 0.0 and clib.exp_det(0.0) == 1.0


def test_sampler_statistics_and_logp():
    rng = np.random.default_rng(0)
    for A in (2, 3):
        logits = np.tile(rng.normal(size=(1, A)).astype(np.float32), (200_000, 1))
        act, logp = clib.sample(logits, seed=7, env_gid0=0, step=11)
        p = np.exp(logits[0] - logits[0].max())
        p /= p.sum()
        freq = np.bincount(act, minlength=A) / len(act)
        np.testing.assert_allclose(freq, p, atol=5e-3)
        np.testing.assert_allclose(logp, np.log(p)[act], atol=2e-6)
    # stream is keyed by (seed, env, step): changing any of them changes the draws
    lg = np.zeros((4096, 2), np.float32)
    base = clib.sample(lg, 7, 0, 11)[0]
    assert (clib.sample(lg, 8, 0, 11)[0] != base).any() and (clib.sample(lg, 7, 0, 12)[0] != base).any()
    assert np.array_equal(clib.sample(lg, 7, 100, 11)[0][:-100], base[100:])   # env id is the counter word


@pytest.mark.parametrize("B", [1, 2, 3, 5, 31, 32, 128, 1000, 4096, 65536 + 17])
def test_permutation_is_bijection(B):
    p = clib.permutation(B, seed=1, epoch_ctr=3, rank=0)
    assert np.array_equal(np.sort(p), np.arange(B, dtype=np.uint32))
    if B >= 128:
        q = clib.permutation(B, seed=1, epoch_ctr=4, rank=0)
        r = clib.permutation(B, seed=1, epoch_ctr=3, rank=1)
        assert (p != q).mean() > 0.9 and (p != r).mean() > 0.9
        assert (p != np.arange(B)).mean() > 0.9


def test_permutation_uniformity():
    # position of element 0 over many epochs is roughly uniform over the 4 minibatch slots
    B, slots = 128, np.zeros(4)
    for e in range(2000):
        p = clib.permutation(B, seed=1, epoch_ctr=e, rank=0)
        slots[int(np.nonzero(p == 0)[0][0]) // 32] += 1
    assert (np.abs(slots / 2000 - 0.25) < 0.04).all()


def test_vector_port_is_deterministic_and_learns_something():
    """The N-env CPU port used as the all-threads CPU arm of bench.py: two runs agree bit for bit, parameters move, the
    step accounting is N*T per update."""
    from oracle import ppo_vector_port as vp
    runs = []
    for _ in range(2):
        p = vp.VectorPort("CartPole-v1", num_envs=8, num_steps=16, seed=3)
        p0 = p.params.copy()
        out = [p.update(), p.update()]
        assert p.env_steps == 2 * 8 * 16 and p.adam_step == 32
        assert np.isfinite(out[-1]["loss"]) and np.abs(p.params - p0).max() > 1e-5
        runs.append(p.params.copy())
    assert np.array_equal(runs[0], runs[1])


def test_mountaincar_oracle_matches_an_independent_restatement():
    """MountainCar-v0 (gym 0.21 mountain_car.py, recalled): the C oracle against a straight Python float64 restatement of
    the published update rule, over random actions incl. the inelastic left wall, the goal and the 200-step TimeLimit."""
    import math
    rng = np.random.default_rng(0)
    for trial in range(20):
        pos, vel = float(rng.uniform(-1.2, 0.6)), float(rng.uniform(-0.07, 0.07))
        s = np.array([pos, vel])
        for _ in range(300):
            a = int(rng.integers(0, 3))
            vel += (a - 1) * 0.001 + math.cos(3 * pos) * (-0.0025)
            vel = min(max(vel, -0.07), 0.07)
            pos += vel
            pos = min(max(pos, -1.2), 0.6)
            if pos == -1.2 and vel < 0:
                vel = 0.0
            done = bool(pos >= 0.5 and vel >= 0)
            s, term = clib.mountaincar_step(s, a)
            assert s[0] == pos and s[1] == vel and term == done
            if done:
                break
    env = clib.OracleVecEnv("MountainCar-v0", 3, seed=4)
    obs = env.reset()
    assert obs.shape == (3, 2) and np.all((obs[:, 0] >= -0.6) & (obs[:, 0] <= -0.4)) and np.all(obs[:, 1] == 0)
    lens = []
    for t in range(450):
        _, r, d, info = env.step(np.ones(3, dtype=np.int32))      # action 1 = no push: never reaches the goal
        assert np.all(r == -1.0)
        lens += [int(x) for x in info["final_length"][d.astype(bool)]]
    assert lens and all(x == 200 for x in lens)                   # TimeLimit(200) of the MountainCar-v0 registration


def test_permutation_inverse_property():
    """hypothesis: for arbitrary sizes, seeds, epochs and ranks the keyed Feistel permutation is a bijection and
    perm_position is its inverse (the statistics kernel relies on it to find a sample's minibatch without a gather)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(B=st.integers(1, 3000), seed=st.integers(0, 2**63 - 1), epoch=st.integers(0, 2**31 - 1), rank=st.integers(0, 7))
    def check(B, seed, epoch, rank):
        p = clib.permutation(B, seed=seed, epoch_ctr=epoch, rank=rank)
        assert np.array_equal(np.sort(p), np.arange(B, dtype=np.uint32))
        L = clib.lib()
        for i in (0, B // 3, B - 1):
            x = L.drl_or_perm_index(i, B, seed, epoch, rank)
            assert x == p[i] and L.drl_or_perm_position(int(x), B, seed, epoch, rank) == i

    check()


def test_gae_matches_a_straight_numpy_restatement():
    """The C oracle's GAE (pinned bit-exactly against the reference's own buffers above) against an independent float64
    evaluation of ppo.py:144-151 on random inputs: agreement to fp32 rounding for arbitrary shapes."""
    rng = np.random.default_rng(5)
    for T, N in [(1, 1), (7, 3), (128, 17)]:
        rew = rng.uniform(0, 1, size=(T + 1, N)).astype(np.float32)
        done = (rng.uniform(size=(T + 1, N)) < 0.1).astype(np.float32)
        val = rng.normal(size=(T + 1, N)).astype(np.float32)
        adv, ret = clib.gae(rew, done, val, 0.99, 0.95)
        want = np.zeros((T + 1, N))
        last = np.zeros(N)
        for t in reversed(range(T)):
            nonterminal = 1.0 - done[t + 1].astype(np.float64)
            delta = rew[t + 1] + 0.99 * val[t + 1].astype(np.float64) * nonterminal - val[t]
            last = delta + 0.99 * 0.95 * nonterminal * last
            want[t] = last
        np.testing.assert_allclose(adv[:T], want[:T], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(ret[:T], want[:T] + val[:T], rtol=2e-5, atol=2e-5)
