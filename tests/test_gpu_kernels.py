"""GPU parity tests: every kernel behind the C ABI vs the CPU oracle / the reference's golden vectors.

Tolerances (north_star): env dynamics <= 1e-6 per step; GAE/returns <= 1e-5 (we assert bit-exact);
minibatch indices and sampled actions bit-exact for a given Philox stream; losses/grads fp32 tolerance.
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import clib, ppo_oracle as po  # noqa: E402


def _dev():
    return torch.device("cuda:0")


def _lib():
    from deep_rl_b200 import _lib as L
    return L


def _net(O, A, H=64):
    L = _lib()
    return L.NetT(O, H, A, 4 if O <= 4 else 8)


def _pack(params_np, O, A, H=64):
    L = _lib()
    net = _net(O, A, H)
    p = torch.tensor(np.asarray(params_np, np.float32), device=_dev())
    packed = torch.zeros(int(L.lib().drl_packed_count(C.byref(net))), dtype=torch.float32, device=_dev())
    L.check(L.lib().drl_pack_params(C.byref(net), p.data_ptr(), packed.data_ptr(), L.stream_ptr()))
    return net, p, packed


def _rand_params(O, A, seed, scale=1.0, H=64):
    return (po.init_params(O, H, A, seed) * scale + (0.05 if H == 64 else 0.02) * torch.randn(po.param_count(O, H, A))).numpy()


# ------------------------------------------------------------------------------------------------
# sampler + permutation: bit-exact
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("A", [2, 3])
def test_sampler_bit_exact(A):
    L = _lib()
    rng = np.random.default_rng(A)
    n = 200_000
    logits = (rng.normal(size=(n, A)) * rng.choice([0.01, 1.0, 10.0, 60.0], size=(n, 1))).astype(np.float32)
    lg = torch.tensor(logits, device=_dev())
    act = torch.empty(n, dtype=torch.int32, device=_dev())
    logp = torch.empty(n, dtype=torch.float32, device=_dev())
    for seed, gid0, step in [(1, 0, 0), (0xDEADBEEFCAFE, 12345, (1 << 33) + 5)]:
        L.check(L.lib().drl_sample(lg.data_ptr(), n, A, seed, gid0, step, act.data_ptr(), logp.data_ptr(), L.stream_ptr()))
        want_a, want_lp = clib.sample(logits, seed, gid0, step)
        assert np.array_equal(act.cpu().numpy(), want_a)
        np.testing.assert_allclose(logp.cpu().numpy(), want_lp, rtol=0, atol=2e-6)


@pytest.mark.parametrize("B", [1, 2, 3, 5, 128, 1000, 4096 * 128])
def test_permutation_bit_exact(B):
    L = _lib()
    idx = torch.empty(B, dtype=torch.int32, device=_dev())
    for seed, ctr, rank in [(1, 0, 0), (99, 7, 3)]:
        L.check(L.lib().drl_permutation(idx.data_ptr(), B, seed, ctr, rank, L.stream_ptr()))
        got = idx.cpu().numpy().view(np.uint32)
        if B <= 200_000:
            assert np.array_equal(got, clib.permutation(B, seed, ctr, rank))
        else:
            head = np.array([clib.lib().drl_or_perm_index(i, B, seed, ctr, rank) for i in range(4096)], dtype=np.uint32)
            assert np.array_equal(got[:4096], head)
            assert np.array_equal(np.sort(got), np.arange(B, dtype=np.uint32))      # size-independent property


# ------------------------------------------------------------------------------------------------
# env kernels
# ------------------------------------------------------------------------------------------------
def _action_seq(kind, n, steps, A):
    rng = np.random.default_rng(7)
    seqs = np.zeros((steps, n), np.int32)
    seqs[:, 0] = 0
    seqs[:, 1] = A - 1
    seqs[:, 2] = np.arange(steps) % 2 * (A - 1)
    seqs[:, 3:] = rng.integers(0, A, size=(steps, n - 3))
    return seqs


@pytest.mark.parametrize("env_id", ["CartPole-v1", "Acrobot-v1", "MountainCar-v0"])
def test_env_step_free_running_vs_oracle(env_id):
    """>= 600 steps on fixed action sequences (crosses the 500-step TimeLimit; 200 for MountainCar), auto-reset included.
    CartPole runs free for all 640 steps.  Acrobot is a chaotic double pendulum (1-ulp sin/cos differences grow
    exponentially), so its float64 state is re-synchronised from the oracle every 64 steps; the per-step bar of
    1e-6 is unchanged."""
    import deep_rl_b200 as drl
    n, steps = 64, 640
    env = drl.make(env_id, num_envs=n, seed=11, env_gid0=1000)
    ora = clib.OracleVecEnv(env_id, n, seed=11, env_gid0=1000)
    obs = env.reset()
    np.testing.assert_allclose(obs.cpu().numpy(), ora.reset(), rtol=0, atol=1e-7)
    np.testing.assert_allclose(env.get_state().cpu().numpy(), ora.state, rtol=0, atol=1e-15)
    acts = _action_seq(env.kind, n, steps, env.num_actions)
    n_done = 0
    worst_state = 0.0
    for t in range(steps):
        o, r, d, _ = env.step(torch.tensor(acts[t], device=_dev()))
        oo, orr, od, info = ora.step(acts[t])
        assert np.array_equal(d.cpu().numpy(), od.astype(bool)), f"done mismatch at step {t}"
        assert np.array_equal(r.cpu().numpy(), orr)
        np.testing.assert_allclose(o.cpu().numpy(), oo, rtol=0, atol=1e-6, err_msg=f"step {t}")
        n_done += int(od.sum())
        if env_id == "Acrobot-v1" and t % 64 == 63:
            worst_state = max(worst_state, float(np.abs(env.get_state().cpu().numpy() - ora.state).max()))
            env.set_state(torch.tensor(ora.state))
    np.testing.assert_allclose(env.get_state().cpu().numpy(), ora.state, rtol=0, atol=1e-9)
    assert worst_state < 1e-7
    assert np.array_equal(env.elapsed.cpu().numpy(), ora.elapsed)
    assert np.array_equal(env.ep_len.cpu().numpy(), ora.ep_len)
    np.testing.assert_allclose(env.ep_ret.cpu().numpy(), ora.ep_ret, rtol=0, atol=0)
    cnt, sum_ret, sum_len, entries = env.log.drain()
    assert cnt == n_done and n_done >= n and len(entries) == n_done


def test_env_teacher_forced_known_answers():
    import deep_rl_b200 as drl
    env = drl.make("CartPole-v1", num_envs=2, seed=1)
    env.reset()
    env.set_state(torch.tensor([[0, 0, 0, 0], [0.01, -0.02, 0.03, 0.04]], dtype=torch.float64))
    for a in [(1, 0), (1, 1), (0, 1)]:
        env.step(torch.tensor(a))
    st = env.get_state().cpu().numpy()
    np.testing.assert_allclose(st[0], (0.011707317073170733, 0.19520443149030264, -0.017560975609756099, -0.29453262531977781), atol=1e-15)
    np.testing.assert_allclose(st[1], (0.0048718496976496205, 0.1736941202746195, 0.038823538610649126, -0.22139198924408143), atol=1e-15)
    env = drl.make("Acrobot-v1", num_envs=1, seed=1)
    env.reset()
    env.set_state(torch.zeros((1, 4), dtype=torch.float64))
    for a in [2, 2, 0, 1]:
        obs, r, d, _ = env.step(torch.tensor([a]))
        assert r == -1.0 and not d
    np.testing.assert_allclose(env.get_state().cpu().numpy()[0],
                               (-0.049859025656954617, 0.15938587589964667, 0.13865703759659068, -0.24781201811290368), atol=1e-13)


def test_single_env_gym_api_episode_info():
    import deep_rl_b200 as drl
    env = drl.make("CartPole-v1", num_envs=1, seed=1)
    obs = env.reset()
    assert obs.shape == (4,) and obs.dtype == torch.float32
    ret, n = 0.0, 0
    for t in range(300):
        obs, r, d, info = env.step(torch.tensor(1))
        assert isinstance(r, float) and isinstance(d, bool)
        ret += r
        n += 1
        if d:
            assert info["episode"]["r"] == ret and info["episode"]["l"] == n
            break
    assert d and n < 30


# ------------------------------------------------------------------------------------------------
# model forward
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("O,A", [(4, 2), (6, 3), (2, 3)])
@pytest.mark.parametrize("n", [1, 7, 8, 1000, 65537])
def test_policy_forward_vs_torch(O, A, n):
    L = _lib()
    torch.manual_seed(n)
    params = _rand_params(O, A, seed=3)
    net, p, packed = _pack(params, O, A)
    OP = net.obs_stride
    obs = torch.randn(n, OP) * 1.5
    obs[:, O:] = 0
    od = obs.to(_dev())
    logits = torch.empty((n, A), dtype=torch.float32, device=_dev())
    value = torch.empty(n, dtype=torch.float32, device=_dev())
    L.check(L.lib().drl_policy_forward(C.byref(net), packed.data_ptr(), od.data_ptr(), n, logits.data_ptr(), value.data_ptr(), L.stream_ptr()))
    wl, wv = po.mlp_forward(torch.tensor(params), obs[:, :O], O, 64, A)
    np.testing.assert_allclose(logits.cpu().numpy(), wl.numpy(), rtol=0, atol=1e-5)
    np.testing.assert_allclose(value.cpu().numpy(), wv.numpy(), rtol=0, atol=1e-5)


def test_agent_surface_matches_reference_names():
    import deep_rl_b200 as drl
    env = drl.make("CartPole-v1", num_envs=4, seed=1)
    torch.manual_seed(1)
    agent = drl.ActorCritic(env)
    assert list(agent.state_dict().keys()) == po.PARAM_NAMES
    # same init draws as the reference's ActorCritic under torch.manual_seed(1)
    np.testing.assert_array_equal(agent.flat_params.cpu().numpy(), po.init_params(4, 64, 2, 1).numpy())
    obs = env.reset()
    v = agent.get_value(obs)
    dist = agent.get_action_distribution(obs)
    a, lp = agent.get_action(obs)
    a2, lp2, ent, v2 = agent.get_action_and_value(obs, a)
    assert v.shape == (4,) and a.shape == (4,) and a.dtype == torch.int64
    wl, wv = po.mlp_forward(agent.flat_params.cpu(), obs.cpu(), 4, 64, 2)
    np.testing.assert_allclose(v.cpu().numpy(), wv.numpy(), atol=1e-5)
    np.testing.assert_allclose(dist.logits.cpu().numpy(), torch.log_softmax(wl, -1).numpy(), atol=1e-5)
    np.testing.assert_allclose(lp.cpu().numpy(), dist.log_prob(a).cpu().numpy(), atol=1e-5)
    np.testing.assert_allclose(lp2.cpu().numpy(), lp.cpu().numpy(), atol=1e-5)
    np.testing.assert_allclose(v2.cpu().numpy(), v.cpu().numpy(), atol=0)
    # torch-side edit of the parameters is picked up after sync()
    with torch.no_grad():
        agent.critic[4].bias.add_(1.0)
    agent.sync()
    np.testing.assert_allclose(agent.get_value(obs).cpu().numpy(), wv.numpy() + 1.0, atol=1e-5)
    single = agent.get_value(obs[0])
    assert single.shape == ()


# ------------------------------------------------------------------------------------------------
# GAE
# ------------------------------------------------------------------------------------------------
def _gae_gpu(rew, done, val, gamma, lam, obs=None, act=None, logp=None, O=4, A=2):
    L = _lib()
    net = _net(O, A)
    T1, N = rew.shape
    d = _dev()
    t_rew = torch.tensor(rew, device=d)
    t_done = torch.tensor(done.astype(np.uint8), device=d)
    t_val = torch.tensor(val, device=d)
    adv = torch.full((T1, N), 7.0, dtype=torch.float32, device=d)
    ret = torch.full((T1, N), 7.0, dtype=torch.float32, device=d)
    rec = None
    t_obs = t_act = t_logp = None
    if obs is not None:
        t_obs = torch.tensor(obs, device=d)
        t_act = torch.tensor(act.astype(np.uint8), device=d)
        t_logp = torch.tensor(logp, device=d)
        rec = torch.zeros(((T1 - 1) * N, L.lib().drl_record_width(C.byref(net))), dtype=torch.float32, device=d)
    buf = L.RolloutBufT(L.ptr(t_obs), L.ptr(t_act), L.ptr(t_logp), t_val.data_ptr(), t_rew.data_ptr(), t_done.data_ptr())
    L.check(L.lib().drl_gae(C.byref(buf), C.byref(net), T1 - 1, N, gamma, lam, adv.data_ptr(), ret.data_ptr(), L.ptr(rec), L.stream_ptr()))
    return adv.cpu().numpy(), ret.cpu().numpy(), None if rec is None else rec.cpu().numpy()


@pytest.mark.parametrize("upd", [0, 1, 2, 50, 155])
def test_gae_vs_reference_golden(golden, upd):
    adv, ret, _ = _gae_gpu(golden[f"u{upd}_rewards"][:, None], golden[f"u{upd}_dones"][:, None],
                           golden[f"u{upd}_values"][:, None], 0.99, 0.95)
    assert np.array_equal(adv[:, 0], golden[f"u{upd}_advantages"])        # bit-exact (bar: 1e-5)
    assert np.array_equal(ret[:, 0], golden[f"u{upd}_returns"])


@pytest.mark.parametrize("T,N,p_done,O,A", [(128, 4096, 0.05, 4, 2), (128, 4096, 1 / 500, 4, 2), (256, 777, 0.05, 6, 3), (1, 1, 0.5, 4, 2)])
def test_gae_and_records_vs_oracle(T, N, p_done, O, A):
    rng = np.random.default_rng(0)
    OP = 4 if O <= 4 else 8
    val = rng.normal(size=(T + 1, N)).astype(np.float32)
    rew = rng.uniform(size=(T + 1, N)).astype(np.float32)
    done = (rng.uniform(size=(T + 1, N)) < p_done).astype(np.float32)
    obs = rng.normal(size=(T + 1, N, OP)).astype(np.float32)
    obs[..., O:] = 0
    act = rng.integers(0, A, size=(T + 1, N))
    logp = -rng.uniform(size=(T + 1, N)).astype(np.float32)
    adv, ret, rec = _gae_gpu(rew, done, val, 0.99, 0.95, obs, act, logp, O, A)
    wadv, wret = clib.gae(rew, done, val, 0.99, 0.95)
    assert np.array_equal(adv, wadv) and np.array_equal(ret, wret)
    RW = rec.shape[1]
    B = T * N
    assert np.array_equal(rec[:, :O], obs[:T].reshape(B, OP)[:, :O])
    assert np.array_equal(rec[:, RW - 4], logp[:T].reshape(B))
    assert np.array_equal(rec[:, RW - 3], wadv[:T].reshape(B))
    assert np.array_equal(rec[:, RW - 2], val[:T].reshape(B))
    assert np.array_equal(rec[:, RW - 1].view(np.int32), act[:T].reshape(B).astype(np.int32))


# ------------------------------------------------------------------------------------------------
# minibatch loss + backward, statistics, clip + Adam
# ------------------------------------------------------------------------------------------------
def _make_records(obs, act, logp, adv, val, O, RW):
    B = obs.shape[0]
    rec = np.zeros((B, RW), np.float32)
    rec[:, :O] = obs[:, :O]
    rec[:, RW - 4] = logp
    rec[:, RW - 3] = adv
    rec[:, RW - 2] = val
    rec[:, RW - 1] = np.asarray(act, np.int32).view(np.float32)
    return rec


def _grad_gpu(params, rec, idx, start, count, O, A, coef=(0.2, 0.01, 0.5), stats=None, flags=0, H=64):
    L = _lib()
    d = _dev()
    net, p, packed = _pack(params, O, A, H)
    t_rec = torch.tensor(rec, device=d)
    t_idx = None if idx is None else torch.tensor(np.asarray(idx, np.int64).astype(np.int32), device=d)
    B = rec.shape[0]
    ws_bytes = int(L.lib().drl_workspace_bytes(C.byref(net)))
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=d)
    n_idx = B if idx is None else len(idx)
    if stats is None:
        st = torch.zeros((16, 2), dtype=torch.float32, device=d)
        L.check(L.lib().drl_adv_stats(C.byref(net), t_rec.data_ptr(), L.ptr(t_idx), n_idx, count, st.data_ptr(), ws.data_ptr(), ws_bytes, L.stream_ptr()))
        assert start % count == 0
        st_ptr = st.data_ptr() + 8 * (start // count)
    else:
        st = torch.tensor(np.asarray(stats, np.float32), device=d)
        st_ptr = st.data_ptr()
    grad = torch.zeros(p.numel(), dtype=torch.float32, device=d)
    terms = torch.zeros(8, dtype=torch.float32, device=d)
    cf = L.PpoCoefT(*coef)
    L.check(L.lib().drl_ppo_minibatch_grad(C.byref(net), packed.data_ptr(), t_rec.data_ptr(), L.ptr(t_idx), start, count, st_ptr,
                                           C.byref(cf), grad.data_ptr(), terms.data_ptr(), ws.data_ptr(), ws_bytes, flags, L.stream_ptr()))
    return grad.cpu().numpy(), terms.cpu().numpy(), st.cpu().numpy()


def test_minibatch_grad_vs_reference_golden(golden):
    g = golden
    rec = _make_records(g["u0_observations"][:128], g["u0_actions"][:128], g["u0_log_probs"][:128],
                        g["u0_advantages"][:128], g["u0_values"][:128], 4, 8)
    params = g["init_params"]
    for i in range(16):
        perm = g["perms"][i // 4]
        grad, terms, _ = _grad_gpu(params, rec, perm, (i % 4) * 32, 32, 4, 2)
        np.testing.assert_allclose(terms[:4], g["mb_terms"][i], rtol=2e-6, atol=2e-6, err_msg=f"minibatch {i}")
        np.testing.assert_allclose(grad, g["mb_grad_pre"][i], rtol=1e-4, atol=2e-6, err_msg=f"minibatch {i}")
        params = g["mb_params_after"][i]


@pytest.mark.parametrize("O,A,B,M", [(4, 2, 4096, 1024), (6, 3, 3000, 750), (4, 2, 100_000, 25_000), (4, 2, 70, 70), (2, 3, 2000, 500)])
def test_minibatch_grad_vs_torch_oracle(O, A, B, M):
    rng = np.random.default_rng(B)
    torch.manual_seed(B)
    RW = 8 if O <= 4 else 16
    params = _rand_params(O, A, seed=5)
    obs = rng.normal(size=(B, O)).astype(np.float32)
    act = rng.integers(0, A, size=B)
    logits, v = po.mlp_forward(torch.tensor(params), torch.tensor(obs), O, 64, A)
    logp_all = torch.log_softmax(logits, -1).numpy()
    logp_old = (logp_all[np.arange(B), act] + rng.normal(scale=0.15, size=B)).astype(np.float32)   # ratios on both sides of the clip
    val_old = (v.numpy() + rng.normal(scale=0.3, size=B)).astype(np.float32)
    adv = rng.normal(size=B).astype(np.float32) * 2 + 0.3
    ret = adv + val_old
    rec = _make_records(obs, act, logp_old, adv, val_old, O, RW)
    idx = clib.permutation(B, 3, 1, 0)
    for k in range(min(2, B // M)):
        sel = idx[k * M:(k + 1) * M].astype(np.int64)
        grad, terms, st = _grad_gpu(params, rec, idx, k * M, M, O, A)
        wt, wg = po.minibatch_loss_and_grad(params, obs[sel], act[sel], logp_old[sel], adv[sel], ret[sel], val_old[sel], O, 64, A)
        a64 = adv[sel].astype(np.float64)
        np.testing.assert_allclose(st[k], [a64.mean(), a64.std(ddof=1)], rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(terms[:4], wt, rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(grad, wg, rtol=2e-4, atol=2e-6)
        assert 0.05 < terms[5] < 0.95      # clip fraction: both branches exercised


@pytest.mark.parametrize("B,M", [(128, 32), (4096 * 128, 131072), (1000, 250), (130, 32), (7, 7)])
def test_adv_stats_perm_matches_gather_and_oracle(B, M):
    """Statistics via the inverse permutation (no gather) == statistics via the materialised permutation."""
    L = _lib()
    d = _dev()
    net = _net(4, 2)
    rng = np.random.default_rng(B)
    adv = (rng.normal(size=B) * 3 + 1).astype(np.float32)
    t_adv = torch.tensor(adv, device=d)
    ws_bytes = int(L.lib().drl_workspace_bytes(C.byref(net)))
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=d)
    st = torch.zeros((16, 2), dtype=torch.float32, device=d)
    seed, ctr, rank = 5, 9, 2
    L.check(L.lib().drl_adv_stats_perm(C.byref(net), t_adv.data_ptr(), B, M, seed, ctr, rank, st.data_ptr(), ws.data_ptr(), ws_bytes, L.stream_ptr()))
    got = st.cpu().numpy()
    idx = torch.empty(B, dtype=torch.int32, device=d)
    L.check(L.lib().drl_permutation(idx.data_ptr(), B, seed, ctr, rank, L.stream_ptr()))
    perm = idx.cpu().numpy().view(np.uint32).astype(np.int64)
    if B <= 200_000:
        assert np.array_equal(perm, clib.permutation(B, seed, ctr, rank))
    nmb = (B + M - 1) // M
    for k in range(nmb):
        a64 = adv[perm[k * M:(k + 1) * M]].astype(np.float64)
        want = [a64.mean(), a64.std(ddof=1) if len(a64) > 1 else np.nan]
        if len(a64) > 1:
            np.testing.assert_allclose(got[k], want, rtol=2e-6, atol=1e-7)


def _rel_l2(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def test_tc_minibatch_grad_vs_reference_golden(golden):
    """tcgen05 path (bf16 operands, fp32 accumulation) against the reference's own gradients: bf16 tolerance."""
    g = golden
    rec = _make_records(g["u0_observations"][:128], g["u0_actions"][:128], g["u0_log_probs"][:128],
                        g["u0_advantages"][:128], g["u0_values"][:128], 4, 8)
    params = g["init_params"]
    for i in range(16):
        perm = g["perms"][i // 4]
        grad, terms, _ = _grad_gpu(params, rec, perm, (i % 4) * 32, 32, 4, 2, flags=1)
        np.testing.assert_allclose(terms[:4], g["mb_terms"][i], rtol=2e-2, atol=5e-3, err_msg=f"minibatch {i}")
        assert _rel_l2(grad, g["mb_grad_pre"][i]) < 0.1, (i, _rel_l2(grad, g["mb_grad_pre"][i]))
        params = g["mb_params_after"][i]


@pytest.mark.parametrize("O,A,B,M", [(4, 2, 4096, 1024), (6, 3, 3000, 750), (4, 2, 100_000, 25_000), (4, 2, 70, 70), (4, 2, 129, 129), (2, 3, 2000, 500)])
def test_tc_minibatch_grad_vs_torch_oracle(O, A, B, M):
    rng = np.random.default_rng(B)
    RW = 8 if O <= 4 else 16
    params = _rand_params(O, A, seed=5)
    obs = rng.normal(size=(B, O)).astype(np.float32)
    act = rng.integers(0, A, size=B)
    logits, v = po.mlp_forward(torch.tensor(params), torch.tensor(obs), O, 64, A)
    logp_all = torch.log_softmax(logits, -1).numpy()
    logp_old = (logp_all[np.arange(B), act] + rng.normal(scale=0.15, size=B)).astype(np.float32)
    val_old = (v.numpy() + rng.normal(scale=0.3, size=B)).astype(np.float32)
    adv = rng.normal(size=B).astype(np.float32) * 2 + 0.3
    ret = adv + val_old
    rec = _make_records(obs, act, logp_old, adv, val_old, O, RW)
    idx = clib.permutation(B, 3, 1, 0)
    for k in range(min(2, B // M)):
        sel = idx[k * M:(k + 1) * M].astype(np.int64)
        grad, terms, st = _grad_gpu(params, rec, idx, k * M, M, O, A, flags=1)
        wt, wg = po.minibatch_loss_and_grad(params, obs[sel], act[sel], logp_old[sel], adv[sel], ret[sel], val_old[sel], O, 64, A)
        g32, t32, _ = _grad_gpu(params, rec, idx, k * M, M, O, A, flags=0)
        print(f"O={O} B={B} mb={k}: rel L2 grad err tc {_rel_l2(grad, wg):.2e} (fp32 path {_rel_l2(g32, wg):.2e}); terms tc {terms[:4]} want {wt}")
        np.testing.assert_allclose(terms[:4], wt, rtol=2e-2, atol=5e-3)
        # bf16 operand rounding (2^-9 per element, unbiased) against sums with heavy cancellation (normalised
        # advantages): measured 0.2-6 % relative L2 error of the whole gradient, up to ~10 % on a single bias
        # vector; a layout / indexing bug gives O(1).
        assert _rel_l2(grad, wg) < 0.1
        off = 0
        for name, shp in zip(po.PARAM_NAMES, po.param_shapes(O, 64, A)):   # a wrong block would hide in the global norm
            n = int(np.prod(shp))
            e = float(np.linalg.norm(grad[off:off + n] - wg[off:off + n]) / max(np.linalg.norm(wg[off:off + n]), 0.05 * np.linalg.norm(wg)))
            assert e < 0.25, (name, e)
            off += n
        assert abs(terms[5] - t32[5]) < 0.02       # clip fraction agrees with the fp32 path


def _gpu_tanh(x: torch.Tensor) -> torch.Tensor:
    """The device's tanh.approx.f32 (activation of the tensor-core kernels), for the bf16-emulating oracle."""
    L = _lib()
    xd = x.detach().to(_dev(), torch.float32).contiguous()
    yd = torch.empty_like(xd)
    L.check(L.lib().drl_selftest_tanh(xd.data_ptr(), yd.data_ptr(), xd.numel(), L.stream_ptr()))
    return yd.cpu().reshape(x.shape)


def _assert_per_tensor(grad, want, O, A, tol, floor=1e-3, H=64):
    """Relative L2 error of every one of the twelve gradient tensors (a wrong block would hide in the global norm); the
    denominator is floored at `floor` x the whole-gradient norm so that a near-zero tensor cannot blow the ratio up."""
    off, worst = 0, ("", 0.0)
    total = float(np.linalg.norm(want))
    for name, shp in zip(po.PARAM_NAMES, po.param_shapes(O, H, A)):
        n = int(np.prod(shp))
        e = float(np.linalg.norm(grad[off:off + n] - want[off:off + n]) / max(np.linalg.norm(want[off:off + n]), floor * total))
        assert e < tol, (name, e)
        worst = max(worst, (name, e), key=lambda t: t[1])
        off += n
    return worst


def test_selftest_tanh_is_the_sfu_instruction():
    x = torch.linspace(-6, 6, 100_001)
    y = _gpu_tanh(x)
    assert float((y - torch.tanh(x)).abs().max()) < 1.5e-3          # tanh.approx.f32: ~2^-11 relative
    assert torch.equal(y, _gpu_tanh(x)) and float(y.abs().max()) <= 1.0


@pytest.mark.parametrize("O,A,B,M", [(4, 2, 4096, 1024), (6, 3, 3000, 750), (4, 2, 100_000, 25_000), (4, 2, 70, 70), (4, 2, 129, 129), (2, 3, 2000, 500)])
def test_tc_minibatch_grad_vs_bf16_emulating_oracle(O, A, B, M):
    """ppo_grad_tc_kernel against the oracle that rounds the GEMM operands to bf16 at the kernel's own rounding points and uses
    the device's tanh.approx: every gradient tensor within 2e-3 relative L2, loss terms within 1e-4 -- 50x tighter than the
    distance to the fp32 maths (test_tc_minibatch_grad_vs_torch_oracle), so an off-by-20 % bias block cannot hide."""
    rng = np.random.default_rng(B)
    RW = 8 if O <= 4 else 16
    params = _rand_params(O, A, seed=5)
    obs = rng.normal(size=(B, O)).astype(np.float32)
    act = rng.integers(0, A, size=B)
    logits, v = po.mlp_forward(torch.tensor(params), torch.tensor(obs), O, 64, A)
    logp_all = torch.log_softmax(logits, -1).numpy()
    logp_old = (logp_all[np.arange(B), act] + rng.normal(scale=0.15, size=B)).astype(np.float32)
    val_old = (v.numpy() + rng.normal(scale=0.3, size=B)).astype(np.float32)
    adv = rng.normal(size=B).astype(np.float32) * 2 + 0.3
    rec = _make_records(obs, act, logp_old, adv, val_old, O, RW)
    idx = clib.permutation(B, 3, 1, 0)
    for k in range(min(2, B // M)):
        sel = idx[k * M:(k + 1) * M].astype(np.int64)
        grad, terms, st = _grad_gpu(params, rec, idx, k * M, M, O, A, flags=1)
        wt, wg = po.minibatch_grad_bf16_emulated(params, obs[sel], act[sel], logp_old[sel], adv[sel], val_old[sel], O, 64, A,
                                                 float(st[k, 0]), float(st[k, 1]), tanh_fn=_gpu_tanh)
        worst = _assert_per_tensor(grad, wg, O, A, tol=2e-3)
        print(f"O={O} B={B} mb={k}: whole-gradient rel L2 vs emulating oracle {_rel_l2(grad, wg):.2e}, worst tensor {worst}")
        assert _rel_l2(grad, wg) < 1e-3
        np.testing.assert_allclose(terms[:6], wt, rtol=1e-4, atol=2e-5)


def test_tc_minibatch_grad_vs_bf16_emulation_on_reference_golden(golden):
    """Same bar on the reference's own first update (its observations, actions, log-probs, advantages, permutations)."""
    g = golden
    rec = _make_records(g["u0_observations"][:128], g["u0_actions"][:128], g["u0_log_probs"][:128],
                        g["u0_advantages"][:128], g["u0_values"][:128], 4, 8)
    params = g["init_params"]
    for i in range(16):
        perm = g["perms"][i // 4]
        sel = perm[(i % 4) * 32:(i % 4) * 32 + 32]
        grad, terms, st = _grad_gpu(params, rec, perm, (i % 4) * 32, 32, 4, 2, flags=1)
        wt, wg = po.minibatch_grad_bf16_emulated(params, g["u0_observations"][:128][sel], g["u0_actions"][:128][sel],
                                                 g["u0_log_probs"][:128][sel], g["u0_advantages"][:128][sel], g["u0_values"][:128][sel],
                                                 4, 64, 2, float(st[i % 4, 0]), float(st[i % 4, 1]), tanh_fn=_gpu_tanh)
        _assert_per_tensor(grad, wg, 4, 2, tol=2e-3)
        np.testing.assert_allclose(terms[:6], wt, rtol=1e-4, atol=2e-5)
        params = g["mb_params_after"][i]


def test_tc_minibatch_grad_deterministic_and_linear():
    O, A, B = 4, 2, 131_072
    rng = np.random.default_rng(1)
    params = _rand_params(O, A, seed=9)
    rec = _make_records(rng.normal(size=(B, O)).astype(np.float32), rng.integers(0, A, size=B),
                        -rng.uniform(0.3, 1.2, size=B).astype(np.float32), rng.normal(size=B).astype(np.float32),
                        rng.normal(size=B).astype(np.float32), O, 8)
    g1, t1, _ = _grad_gpu(params, rec, None, 0, B, O, A, stats=[0.0, 1.0], flags=1)
    g2, t2, _ = _grad_gpu(params, rec, None, 0, B, O, A, stats=[0.0, 1.0], flags=1)
    assert np.array_equal(g1, g2) and np.array_equal(t1, t2)
    g32, t32, _ = _grad_gpu(params, rec, None, 0, B, O, A, stats=[0.0, 1.0], flags=0)
    assert _rel_l2(g1, g32) < 0.1
    np.testing.assert_allclose(t1[:4], t32[:4], rtol=2e-2, atol=5e-3)


def test_minibatch_grad_linearity_in_coefficients():
    """Size-independent property at full minibatch size (131,072 samples of C2): the gradient is linear in
    (ent_coef, vf_coef) and the result is deterministic run to run."""
    O, A, B = 4, 2, 131_072
    rng = np.random.default_rng(1)
    torch.manual_seed(1)
    params = _rand_params(O, A, seed=9)
    rec = _make_records(rng.normal(size=(B, O)).astype(np.float32), rng.integers(0, A, size=B),
                        -rng.uniform(0.3, 1.2, size=B).astype(np.float32), rng.normal(size=B).astype(np.float32),
                        rng.normal(size=B).astype(np.float32), O, 8)
    g00, _, _ = _grad_gpu(params, rec, None, 0, B, O, A, coef=(0.2, 0.0, 0.0), stats=[0.0, 1.0])
    g10, _, _ = _grad_gpu(params, rec, None, 0, B, O, A, coef=(0.2, 0.01, 0.0), stats=[0.0, 1.0])
    g01, _, _ = _grad_gpu(params, rec, None, 0, B, O, A, coef=(0.2, 0.0, 0.5), stats=[0.0, 1.0])
    g11, _, _ = _grad_gpu(params, rec, None, 0, B, O, A, coef=(0.2, 0.01, 0.5), stats=[0.0, 1.0])
    g11b, _, _ = _grad_gpu(params, rec, None, 0, B, O, A, coef=(0.2, 0.01, 0.5), stats=[0.0, 1.0])
    assert np.array_equal(g11, g11b)
    np.testing.assert_allclose(g11, g10 + g01 - g00, rtol=1e-4, atol=1e-6)
    assert np.abs(g01[4610:] - g00[4610:]).max() > 1e-4 and np.abs(g01[:4610] - g00[:4610]).max() == 0   # vf only moves the critic


def test_clip_adam_vs_reference_golden(golden):
    L = _lib()
    g = golden
    d = _dev()
    net = _net(4, 2)
    p = torch.tensor(g["init_params"], device=d)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    packed = torch.zeros(int(L.lib().drl_packed_count(C.byref(net))), dtype=torch.float32, device=d)
    norm = torch.zeros(1, dtype=torch.float32, device=d)
    for i in range(16):
        grad = torch.tensor(g["mb_grad_pre"][i], device=d)
        L.check(L.lib().drl_clip_adam(C.byref(net), p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), i + 1,
                                      float(g["mb_lr"][i]), 0.9, 0.999, 1e-5, 0.5, 1.0, packed.data_ptr(), norm.data_ptr(), L.stream_ptr()))
        assert abs(float(norm.item()) - g["mb_norm"][i]) < 1e-5
        np.testing.assert_allclose(p.cpu().numpy(), g["mb_params_after"][i], rtol=0, atol=3e-7, err_msg=f"step {i}")
        if i in (0, 15):   # fused re-pack equals a fresh pack of the updated parameters
            _, _, fresh = _pack(p.cpu().numpy(), 4, 2)
            assert torch.equal(packed, fresh)
        p.copy_(torch.tensor(g["mb_params_after"][i]))


def test_clip_adam_world_scale_and_acrobot():
    L = _lib()
    d = _dev()
    net = _net(6, 3)
    P = po.param_count(6, 64, 3)
    rng = np.random.default_rng(2)
    params, grad = rng.normal(size=P).astype(np.float32), rng.normal(size=P).astype(np.float32) * 0.01
    m0, v0 = rng.normal(size=P).astype(np.float32) * 0.01, rng.uniform(size=P).astype(np.float32) * 1e-4
    p, gsum, m, v = (torch.tensor(x, device=d) for x in (params, grad * 4, m0, v0))      # 4 ranks summed
    L.check(L.lib().drl_clip_adam(C.byref(net), p.data_ptr(), gsum.data_ptr(), m.data_ptr(), v.data_ptr(), 37, 1e-3, 0.9, 0.999,
                                  1e-5, 0.5, 0.25, 0, 0, L.stream_ptr()))
    wp, wm, wv, _ = po.clip_adam(params, grad, m0, v0, 37, 1e-3)
    np.testing.assert_allclose(p.cpu().numpy(), wp, rtol=0, atol=3e-7)
    np.testing.assert_allclose(m.cpu().numpy(), wm, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(v.cpu().numpy(), wv, rtol=1e-6, atol=1e-12)


# ------------------------------------------------------------------------------------------------
# 256-wide actor-critic (BASELINE config C5): tensor-core path only
# ------------------------------------------------------------------------------------------------
def test_h256_shapes():
    L = _lib()
    for O, A in ((4, 2), (6, 3), (2, 3)):
        net = _net(O, A, 256)
        assert L.lib().drl_param_count(C.byref(net)) == po.param_count(O, 256, A)
    assert L.lib().drl_param_count(C.byref(_net(4, 2, 256))) == 134_915
    assert L.lib().drl_param_count(C.byref(_net(4, 2, 128))) == -1


@pytest.mark.parametrize("O,A", [(4, 2), (6, 3), (2, 3)])
@pytest.mark.parametrize("n", [1, 127, 128, 129, 5000, 40_000])
def test_policy_forward_h256(O, A, n):
    """drl_policy_forward at hidden = 256 (mlp256_kernel, forward mode): <= 5e-4 from the bf16-emulating oracle with the
    device's tanh.approx, <= 3e-2 from the fp32 torch maths (deep_rl/ppo.py:49-54)."""
    L = _lib()
    torch.manual_seed(n)
    params = _rand_params(O, A, seed=3, H=256)
    net, p, packed = _pack(params, O, A, 256)
    OP = net.obs_stride
    obs = torch.randn(n, OP) * 1.5
    obs[:, O:] = 0
    od = obs.to(_dev())
    logits = torch.full((n, A), float("nan"), dtype=torch.float32, device=_dev())
    value = torch.full((n,), float("nan"), dtype=torch.float32, device=_dev())
    L.check(L.lib().drl_policy_forward(C.byref(net), packed.data_ptr(), od.data_ptr(), n, logits.data_ptr(), value.data_ptr(), L.stream_ptr()))
    el, ev = po.mlp_forward_bf16_emulated(params, obs[:, :O].numpy(), O, 256, A, tanh_fn=_gpu_tanh)
    # a bf16 rounding of one h1 / h2 element that flips (accumulation-order noise at a rounding boundary) moves an output by up
    # to ~1e-3; that happens to about one sample in 10^4, everything else agrees to fp32 rounding
    for got, want in ((logits.cpu().numpy(), el.numpy()), (value.cpu().numpy(), ev.numpy())):
        err = np.abs(got - want)
        assert err.max() < 3e-3 and np.mean(err > 5e-4) < 2e-3 and np.median(err) < 2e-6, (err.max(), np.mean(err > 5e-4), np.median(err))
    wl, wv = po.mlp_forward(torch.tensor(params), obs[:, :O], O, 256, A)
    np.testing.assert_allclose(logits.cpu().numpy(), wl.numpy(), rtol=0, atol=3e-2)
    np.testing.assert_allclose(value.cpu().numpy(), wv.numpy(), rtol=0, atol=3e-2)


@pytest.mark.parametrize("O,A,B,M", [(4, 2, 4096, 1024), (4, 2, 70, 70), (4, 2, 129, 129), (6, 3, 3000, 750), (2, 3, 2000, 500), (4, 2, 60_000, 30_000)])
def test_minibatch_grad_h256_vs_oracles(O, A, B, M):
    """mlp256_kernel + dw2_gemm256_kernel against the bf16-emulating oracle (every tensor <= 2e-3 relative L2) and against
    torch autograd in fp32 (bf16 tolerance)."""
    H = 256
    rng = np.random.default_rng(B)
    RW = 8 if O <= 4 else 16
    params = _rand_params(O, A, seed=5, H=H)
    obs = rng.normal(size=(B, O)).astype(np.float32)
    act = rng.integers(0, A, size=B)
    logits, v = po.mlp_forward(torch.tensor(params), torch.tensor(obs), O, H, A)
    logp_all = torch.log_softmax(logits, -1).numpy()
    logp_old = (logp_all[np.arange(B), act] + rng.normal(scale=0.15, size=B)).astype(np.float32)
    val_old = (v.numpy() + rng.normal(scale=0.3, size=B)).astype(np.float32)
    adv = rng.normal(size=B).astype(np.float32) * 2 + 0.3
    ret = adv + val_old
    rec = _make_records(obs, act, logp_old, adv, val_old, O, RW)
    idx = clib.permutation(B, 3, 1, 0)
    for k in range(min(2, B // M)):
        sel = idx[k * M:(k + 1) * M].astype(np.int64)
        grad, terms, st = _grad_gpu(params, rec, idx, k * M, M, O, A, flags=1, H=H)
        wt, wg = po.minibatch_grad_bf16_emulated(params, obs[sel], act[sel], logp_old[sel], adv[sel], val_old[sel], O, H, A,
                                                 float(st[k, 0]), float(st[k, 1]), tanh_fn=_gpu_tanh)
        print(f"H=256 O={O} B={B} mb={k}: whole-gradient rel L2 vs emulating oracle {_rel_l2(grad, wg):.2e}")
        worst = _assert_per_tensor(grad, wg, O, A, tol=2e-3, H=H)
        print("   worst tensor", worst)
        np.testing.assert_allclose(terms[:6], wt, rtol=1e-4, atol=2e-5)
        ft, fg = po.minibatch_loss_and_grad(params, obs[sel], act[sel], logp_old[sel], adv[sel], ret[sel], val_old[sel], O, H, A)
        np.testing.assert_allclose(terms[:4], ft, rtol=2e-2, atol=5e-3)
        assert _rel_l2(grad, fg) < 0.1
        g2, t2, _ = _grad_gpu(params, rec, idx, k * M, M, O, A, flags=1, H=H)
        assert np.array_equal(grad, g2) and np.array_equal(terms, t2)          # deterministic run to run
