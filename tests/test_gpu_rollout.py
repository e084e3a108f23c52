"""GPU parity tests of the fused rollout kernel and of whole PPO updates (deep_rl/ppo.py:105-192)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import clib, ppo_oracle as po  # noqa: E402


def _gpu_tanh(x: torch.Tensor) -> torch.Tensor:
    """tanh.approx.f32 evaluated by the device (the activation of the tensor-core kernels) for the bf16-emulating oracle."""
    from deep_rl_b200 import _lib as L
    xd = x.detach().to("cuda:0", torch.float32).contiguous()
    yd = torch.empty_like(xd)
    L.check(L.lib().drl_selftest_tanh(xd.data_ptr(), yd.data_ptr(), xd.numel(), L.stream_ptr()))
    return yd.cpu().reshape(x.shape)


def _check_rollout(env_id, N, T, seed, sub=None, rounds=2, tc=False, hidden=64):
    """Teacher-forced check of drl_rollout against the oracle:
       * the kernel's own logits (debug plane) fed to the ORACLE sampler must give the kernel's actions EXACTLY and its
         log-probs (SURVEY hard part 2: identical logits + identical Philox counter => identical action),
       * those logits and val[t] must equal the oracle forward on the kernel's stored obs[t]: the fp32 torch oracle for
         the CUDA-core kernel (2e-5); for the tensor-core kernel the bf16-emulating oracle with the device's tanh.approx
         (5e-4: layer 1 runs as fp32 FMAs here and as a GEMM in the oracle, so about one h1 element in 10^4 rounds to the
         neighbouring bf16 value, which moves a value by up to ~2e-4; the fp32 oracle is kept as the loose outer bound, 3e-2),
       * the oracle env driven by the kernel's actions must reproduce obs/rew/done (<= 1e-6 per step),
         auto-resets (Philox reset draws) and the episode log included."""
    import deep_rl_b200 as drl
    if sub is not None:
        os.environ["DRL_ROLLOUT_EPW"] = str(sub)
    try:
        cfg = drl.PPOConfig(env_id=env_id, num_envs=N, num_steps=T, seed=seed, rollout_precision="bf16" if tc else "fp32",
                            update_precision="bf16" if tc else "fp32", debug_logits=True, hidden=hidden)
        tr = drl.PPOTrainer(cfg)
        tol = 3e-2 if tc else 2e-5          # distance to the fp32 reference maths
        O, A = tr.env.obs_dim, tr.env.num_actions
        ora = clib.OracleVecEnv(env_id, N, seed=seed)
        obs0 = ora.reset()
        flat = tr.agent.flat_params.cpu()
        for rnd in range(rounds):
            tr.rollout()
            torch.cuda.synchronize()
            obs = tr.observations.cpu().numpy()[:, :, :O]
            act = tr.actions.cpu().numpy().astype(np.int32)
            logp, val = tr.log_probs.cpu().numpy(), tr.values.cpu().numpy()
            rew, done = tr.rewards.cpu().numpy(), tr.dones.cpu().numpy()
            klog = tr.logits.cpu().numpy()
            np.testing.assert_allclose(obs[0], obs0, rtol=0, atol=1e-7)
            if tc:      # one batched emulated forward over the whole rollout
                el, ev = po.mlp_forward_bf16_emulated(flat.numpy(), obs.reshape(-1, O), O, hidden, A, tanh_fn=_gpu_tanh)
                el, ev = el.numpy().reshape(T + 1, N, A), ev.numpy().reshape(T + 1, N)
                etol = 5e-4 if hidden == 64 else 3e-3       # one flipped bf16 rounding moves a 256-wide output by up to ~1e-3
                np.testing.assert_allclose(val, ev, rtol=0, atol=etol, err_msg="value vs bf16-emulating oracle")
                np.testing.assert_allclose(klog, el[:T], rtol=0, atol=etol, err_msg="logits vs bf16-emulating oracle")
                assert np.mean(np.abs(val - ev) > 2e-5) < 0.02, "more than 2 % of the values off by more than fp32 rounding"
            fin = []
            for t in range(T + 1):
                with torch.no_grad():
                    logits, v = po.mlp_forward(flat, torch.from_numpy(np.ascontiguousarray(obs[t])), O, hidden, A)
                np.testing.assert_allclose(val[t], v.numpy(), rtol=0, atol=tol, err_msg=f"value t={t}")
                if t == T:
                    break
                step = ora.step_count
                np.testing.assert_allclose(klog[t], logits.numpy(), rtol=0, atol=tol, err_msg=f"logits t={t}")
                wa, wlp = clib.sample(klog[t], seed, 0, step)            # oracle sampler on the KERNEL's logits: exact
                assert np.array_equal(wa, act[t]), f"sampled actions differ at t={t}: envs {np.nonzero(wa != act[t])[0]}"
                np.testing.assert_allclose(logp[t], wlp, rtol=0, atol=2e-6, err_msg=f"logp t={t}")
                o, r, d, info = ora.step(act[t])
                np.testing.assert_allclose(obs[t + 1], o, rtol=0, atol=1e-6, err_msg=f"obs t={t + 1}")
                assert np.array_equal(rew[t + 1], r) and np.array_equal(done[t + 1], d), f"rew/done t={t + 1}"
                for i in np.nonzero(d)[0]:
                    fin.append((step, int(i), float(info["final_return"][i]), int(info["final_length"][i])))
            obs0 = obs[T]
            cnt, sum_ret, sum_len, entries = tr.env.log.drain()
            assert cnt == len(fin)
            assert entries == sorted(fin)
            np.testing.assert_allclose(tr.env.get_state().cpu().numpy(), ora.state, rtol=0, atol=1e-9)
        return tr
    finally:
        os.environ.pop("DRL_ROLLOUT_EPW", None)


@pytest.mark.parametrize("env_id,N,T,sub", [
    ("CartPole-v1", 1, 128, None),        # the reference's own shape
    ("CartPole-v1", 13, 40, None),        # ragged: not a multiple of the 8-env tile
    ("CartPole-v1", 256, 64, None),
    ("CartPole-v1", 100, 48, 16),         # envs per warp forced: 2 x 8-env tiles
    ("CartPole-v1", 300, 32, 32),         # 4 x 8-env tiles
    ("CartPole-v1", 77, 32, 8),
    ("CartPole-v1", 4096, 16, None),      # C2 width: picks the 4-env tile
    ("Acrobot-v1", 40, 48, None),
    ("Acrobot-v1", 70, 24, 32),
    ("Acrobot-v1", 50, 24, 8),
    ("MountainCar-v0", 45, 40, None),
    ("MountainCar-v0", 64, 24, 16),
])
def test_rollout_vs_oracle(env_id, N, T, sub):
    _check_rollout(env_id, N, T, seed=3, sub=sub)


@pytest.mark.parametrize("env_id,N,T", [("CartPole-v1", 300, 24), ("CartPole-v1", 128, 40), ("CartPole-v1", 1, 16), ("Acrobot-v1", 130, 16),
                                       ("MountainCar-v0", 96, 24)])
def test_tensor_core_rollout_vs_oracle(env_id, N, T):
    _check_rollout(env_id, N, T, seed=3, tc=True)


@pytest.mark.parametrize("rows", [32, 64, 128])
def test_tensor_core_rollout_every_row_variant(rows, monkeypatch):
    """The launcher picks 32 / 64 / 128 envs per CTA (each env on 4 / 2 / 1 GEMM rows) from the env count; force each
    variant on the same small problem."""
    monkeypatch.setenv("DRL_ROLLOUT_ROWS", str(rows))
    _check_rollout("CartPole-v1", 200, 24, seed=4, tc=True)
    _check_rollout("Acrobot-v1", 70, 16, seed=4, tc=True)
    _check_rollout("MountainCar-v0", 90, 16, seed=4, tc=True, rounds=1)


@pytest.mark.parametrize("rows", [32, 64, 128])
def test_tensor_core_rollout_rows_per_cta(rows):
    """Few envs spread over more CTAs (32 / 64 rows of the 128-row tile used): same results as the oracle."""
    os.environ["DRL_ROLLOUT_ROWS"] = str(rows)
    try:
        _check_rollout("CartPole-v1", 200, 20, seed=5, tc=True, rounds=1)
    finally:
        os.environ.pop("DRL_ROLLOUT_ROWS", None)


def test_rollout_world_size_invariance():
    """Global env id feeds the Philox counter: rank 1 of a 2-rank job reproduces envs [N, 2N) of a 1-rank job."""
    import deep_rl_b200 as drl
    T, N = 32, 24
    full = drl.PPOTrainer(drl.PPOConfig(num_envs=2 * N, num_steps=T, seed=5))
    full.rollout()
    half = drl.PPOTrainer(drl.PPOConfig(num_envs=N, num_steps=T, seed=5), rank=1, world=2)
    half.rollout()
    torch.cuda.synchronize()
    for name in ("observations", "actions", "log_probs", "values", "rewards", "dones"):
        a, b = getattr(full, name)[:, N:], getattr(half, name)
        assert torch.equal(a, b), name


def _oracle_update(tr, nu):
    """One whole update on the CPU from the trainer's current state: GAE (C oracle), then the 16 minibatch steps
    with torch autograd + torch.optim.Adam semantics, using the oracle permutation."""
    cfg = tr.cfg
    O, A, N, T = tr.env.obs_dim, tr.env.num_actions, cfg.num_envs, cfg.num_steps
    B, M = cfg.batch_size, cfg.minibatch_size
    obs = tr.observations.cpu().numpy()[:T, :, :O].reshape(B, O)
    act = tr.actions.cpu().numpy()[:T].reshape(B)
    logp = tr.log_probs.cpu().numpy()[:T].reshape(B)
    val = tr.values.cpu().numpy()
    adv, ret = clib.gae(tr.rewards.cpu().numpy(), tr.dones.cpu().numpy().astype(np.float32), val, cfg.gamma, cfg.gae_lambda)
    return obs, act, logp, val[:T].reshape(B), adv, ret, B, M


@pytest.mark.parametrize("env_id,N,T", [("CartPole-v1", 1, 128), ("CartPole-v1", 64, 32), ("Acrobot-v1", 24, 64), ("MountainCar-v0", 32, 48)])
def test_full_update_vs_oracle(env_id, N, T):
    """rollout (GPU) -> [GAE, permutation, statistics, 16 x (loss+backward, clip, Adam)] on GPU vs the CPU oracle
    fed the same rollout buffers."""
    import deep_rl_b200 as drl
    cfg = drl.PPOConfig(env_id=env_id, num_envs=N, num_steps=T, seed=2, total_timesteps=N * T * 10, update_precision="fp32")
    tr = drl.PPOTrainer(cfg)
    nu = cfg.num_updates()
    for upd in range(2):
        params0 = tr.agent.flat_params.cpu().numpy().copy()
        m0, v0 = tr.exp_avg.cpu().numpy().copy(), tr.exp_avg_sq.cpu().numpy().copy()
        step0 = tr.adam_step
        tr.update(nu)
        torch.cuda.synchronize()
        obs, act, logp, val, adv, ret, B, M = _oracle_update(tr, nu)
        assert np.array_equal(tr.advantages.cpu().numpy(), adv) and np.array_equal(tr.returns.cpu().numpy(), ret)
        lr = (1.0 - upd / nu) * cfg.learning_rate
        O, A = tr.env.obs_dim, tr.env.num_actions
        p, m, v, k = params0, m0, v0, step0
        terms_all = tr.loss_terms.cpu().numpy()
        for epoch in range(cfg.update_epochs):
            idx = clib.permutation(B, cfg.seed, upd * cfg.update_epochs + epoch, 0).astype(np.int64)
            for j in range(4):
                sel = idx[j * M:(j + 1) * M]
                terms, grad = po.minibatch_loss_and_grad(p, obs[sel], act[sel], logp[sel], adv[:T].reshape(B)[sel],
                                                         ret[:T].reshape(B)[sel], val[sel], O, 64, A)
                k += 1
                p, m, v, _ = po.clip_adam(p, grad, m, v, k, lr)
                np.testing.assert_allclose(terms_all[epoch * 4 + j, :4], terms, rtol=1e-3, atol=2e-5,
                                           err_msg=f"update {upd} epoch {epoch} mb {j}")
        # Adam normalises the step, so tiny gradient differences can move a weight by O(lr * 1e-3)
        np.testing.assert_allclose(tr.agent.flat_params.cpu().numpy(), p, rtol=0, atol=2e-5)
        assert tr.adam_step == k
    ev = tr.explained_variance()          # ppo.py:194-195 over all T+1 slots, against the float64 restatement
    yv, rv = tr.values.cpu().numpy().ravel(), tr.returns.cpu().numpy().ravel()
    want = 1.0 - np.var((yv - rv).astype(np.float64), ddof=1) / np.var(yv.astype(np.float64), ddof=1)
    assert abs(ev - want) <= 1e-5 * max(1.0, abs(want)), (ev, want)


@pytest.mark.parametrize("env_id,N,T", [("CartPole-v1", 64, 32), ("Acrobot-v1", 24, 64), ("MountainCar-v0", 32, 48)])
def test_full_update_tensor_core_path_tracks_fp32_path(env_id, N, T):
    """Same seeds, one update: the tcgen05 (bf16) update must stay within bf16 tolerance of the fp32 update."""
    import deep_rl_b200 as drl
    out = {}
    for prec in ("fp32", "bf16"):
        cfg = drl.PPOConfig(env_id=env_id, num_envs=N, num_steps=T, seed=2, total_timesteps=N * T * 10, update_precision=prec,
                            rollout_precision="fp32")       # same rollout data for both updates
        tr = drl.PPOTrainer(cfg)
        p0 = tr.agent.flat_params.cpu().numpy().copy()
        tr.update()
        torch.cuda.synchronize()
        out[prec] = (tr.loss_terms.cpu().numpy().copy(), tr.agent.flat_params.cpu().numpy() - p0)
    np.testing.assert_allclose(out["bf16"][0][:, :4], out["fp32"][0][:, :4], rtol=3e-2, atol=1e-2)
    d32, d16 = out["fp32"][1], out["bf16"][1]
    assert np.linalg.norm(d16 - d32) / np.linalg.norm(d32) < 0.15      # 16 Adam-normalised steps
    assert np.isfinite(d16).all()


def test_reference_shape_learning_curve():
    """The reference configuration (1 env, 128 steps, 20k timesteps, seed 1) must learn like the reference does
    (golden: mean return of the first 20 episodes 27.8 -> last 20 episodes 217.1)."""
    import deep_rl_b200 as drl
    best = 0.0
    for seed in (1, 2, 3):
        cfg = drl.PPOConfig(seed=seed)
        tr = drl.PPOTrainer(cfg)
        eps = []
        for _ in range(cfg.num_updates()):
            tr.update()
            eps += [e[2] for e in tr.metrics()["episode_log"]]
        first, last = float(np.mean(eps[:20])), float(np.mean(eps[-20:]))
        assert first < 60
        best = max(best, last)
    assert best > 120, best


def test_many_envs_learns_cartpole():
    import deep_rl_b200 as drl
    cfg = drl.PPOConfig(num_envs=512, num_steps=128, total_timesteps=512 * 128 * 100, seed=1, update_precision="bf16")
    tr = drl.PPOTrainer(cfg)
    assert tr.rollout_flags == 1 and tr.grad_flags == 1      # tcgen05 rollout + update: the benchmarked numerics
    rets = []
    for _ in range(cfg.num_updates()):
        tr.update()
        m = tr.metrics()
        if m["episodes"]:
            rets.append(m["mean_return"])
    assert rets[0] < 40 and max(rets[-5:]) > 150, rets


def test_train_prints_reference_format(capsys):
    import deep_rl_b200 as drl
    drl.train(drl.PPOConfig(total_timesteps=1024, seed=1))
    out = capsys.readouterr().out.strip().splitlines()
    assert len(out) > 10
    import re
    assert all(re.fullmatch(r"global_step=\d+, episodic_return=\d+\.\d\d", ln) for ln in out)
    steps = [int(ln.split(",")[0].split("=")[1]) for ln in out]
    assert steps == sorted(steps) and steps[-1] < 1024


@pytest.mark.parametrize("env_id,N,T,prec", [("Acrobot-v1", 24, 64, "fp32"), ("CartPole-v1", 64, 32, "bf16")])
def test_side_stream_overlap_is_bit_identical(env_id, N, T, prec):
    """Permutations / statistics scheduled on the side stream must not change a single bit of the result (and the result
    must be the same run after run: the kernels have no atomics on the data path)."""
    import deep_rl_b200 as drl
    ref = None
    for overlap in (False, True):
        for _rep in range(6):
            cfg = drl.PPOConfig(env_id=env_id, num_envs=N, num_steps=T, seed=2, total_timesteps=N * T * 10, update_precision=prec,
                                overlap_streams=overlap)
            tr = drl.PPOTrainer(cfg)
            for _ in range(3):
                tr.update(10)
            torch.cuda.synchronize()
            got = (tr.agent.flat_params.cpu().numpy().copy(), tr.loss_terms.cpu().numpy().copy())
            if ref is None:
                ref = got
            assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), f"overlap={overlap}"


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_checkpoint_resume_is_bit_identical(prec):
    """state_dict() after two updates, loaded into a fresh trainer: the next two updates match the uninterrupted run bit for
    bit (model, Adam moments, step counters, env state, Philox counters)."""
    import deep_rl_b200 as drl
    mk = lambda: drl.PPOTrainer(drl.PPOConfig(num_envs=96, num_steps=32, seed=5, total_timesteps=96 * 32 * 10, update_precision=prec))
    a = mk()
    a.update(10); a.update(10)
    sd = a.state_dict()
    a.update(10); a.update(10)
    torch.cuda.synchronize()
    b = mk()
    b.update(10)                      # diverge first: loading must overwrite everything that matters
    b.load_state_dict(sd)
    b.update(10); b.update(10)
    torch.cuda.synchronize()
    assert torch.equal(a.agent.flat_params, b.agent.flat_params)
    assert torch.equal(a.exp_avg, b.exp_avg) and torch.equal(a.exp_avg_sq, b.exp_avg_sq)
    assert torch.equal(a.env.state, b.env.state) and a.env.step_count == b.env.step_count and a.adam_step == b.adam_step
    assert torch.equal(a.loss_terms, b.loss_terms)


def test_checkpoint_mismatch_is_rejected():
    """A checkpoint only resumes bit-identically under the same seed / rank / world / width / precisions: anything else raises."""
    import deep_rl_b200 as drl
    a = drl.PPOTrainer(drl.PPOConfig(num_envs=32, num_steps=16, seed=5))
    a.update(4)
    sd = a.state_dict()
    for kw in (dict(seed=6), dict(update_precision="bf16"), dict(num_envs=64)):
        base = dict(num_envs=32, num_steps=16, seed=5)
        base.update(kw)
        with pytest.raises(ValueError, match="checkpoint does not match"):
            drl.PPOTrainer(drl.PPOConfig(**base)).load_state_dict(sd)
    with pytest.raises(ValueError, match="checkpoint does not match"):
        drl.PPOTrainer(drl.PPOConfig(num_envs=32, num_steps=16, seed=5), rank=1, world=2).load_state_dict(sd)
    drl.PPOTrainer(drl.PPOConfig(num_envs=32, num_steps=16, seed=5)).load_state_dict(sd)


@pytest.mark.parametrize("N,T,want", [(1, 128, "fp32"), (64, 32, "fp32"), (2048, 16, "bf16"), (4096, 8, "bf16")])
def test_default_precision_keeps_ratio_one_at_first_minibatch(N, T, want):
    """Default config: rollout and update share their numerics at every size (fp32 + fp32 below 2048 envs, tcgen05 + tcgen05
    from there), so at epoch 0 / minibatch 0 the re-evaluated log-probs equal the recorded ones: approx_kl ~ 0, clipfrac = 0
    (deep_rl/ppo.py:170-171 -- in the reference the first ratio is exactly 1)."""
    import deep_rl_b200 as drl
    tr = drl.PPOTrainer(drl.PPOConfig(num_envs=N, num_steps=T, seed=3, total_timesteps=N * T * 4))
    assert tr.update_precision == want and tr.rollout_precision == want
    tr.update(4)
    torch.cuda.synchronize()
    terms = tr.loss_terms.cpu().numpy()
    assert abs(terms[0, 4]) < 1e-5 and terms[0, 5] == 0.0, terms[0]
    assert terms[-1, 4] > terms[0, 4]          # later minibatches do move away from the behaviour policy


def test_world_gt_one_without_process_group_raises():
    import deep_rl_b200 as drl
    from deep_rl_b200._lib import DrlError
    tr = drl.PPOTrainer(drl.PPOConfig(num_envs=32, num_steps=16, seed=5), rank=0, world=2)
    tr.rollout()                     # sharded rollouts need no communication
    tr.compute_gae()
    with pytest.raises(DrlError, match="process group"):
        tr.optimize(1e-3)


def test_explained_variance_vs_reference_golden(golden):
    """ppo.py:194-195: the script's own explained_var after its last update, from its own values / returns."""
    from deep_rl_b200 import _lib as L
    d = torch.device("cuda:0")
    y, r = torch.tensor(golden["u155_values"], device=d), torch.tensor(golden["u155_returns"], device=d)
    net = L.NetT(4, 64, 2, 4)
    nb = int(L.lib().drl_workspace_bytes(C.byref(net)))
    ws = torch.zeros(nb, dtype=torch.uint8, device=d)
    out = torch.zeros(1, dtype=torch.float32, device=d)
    L.check(L.lib().drl_explained_variance(y.data_ptr(), r.data_ptr(), y.numel(), out.data_ptr(), ws.data_ptr(), nb, L.stream_ptr()))
    want = float(golden["explained_var"])
    assert abs(float(out.item()) - want) <= 2e-6 * abs(want), (float(out.item()), want)
    # large, multi-CTA case against float64 numpy; constant values -> NaN like the reference
    rng = np.random.default_rng(0)
    yy, rr = rng.normal(size=3_000_001).astype(np.float32) * 3 + 1, rng.normal(size=3_000_001).astype(np.float32)
    ty, trr = torch.tensor(yy, device=d), torch.tensor(rr, device=d)
    for _ in range(2):               # twice: the arrival counter re-arms itself
        L.check(L.lib().drl_explained_variance(ty.data_ptr(), trr.data_ptr(), ty.numel(), out.data_ptr(), ws.data_ptr(), nb, L.stream_ptr()))
        want = 1.0 - np.var((yy - rr).astype(np.float64), ddof=1) / np.var(yy.astype(np.float64), ddof=1)
        assert abs(float(out.item()) - want) < 1e-6
    ty.fill_(2.5)
    L.check(L.lib().drl_explained_variance(ty.data_ptr(), trr.data_ptr(), ty.numel(), out.data_ptr(), ws.data_ptr(), nb, L.stream_ptr()))
    assert np.isnan(float(out.item()))


def test_episode_log_overflow_is_counted():
    """More finished episodes than log entries: the extra ones are dropped AND reported (count and sums stay exact)."""
    import deep_rl_b200 as drl
    tr = drl.PPOTrainer(drl.PPOConfig(num_envs=256, num_steps=64, seed=2))
    tr.env.log = type(tr.env.log)(16, tr.device)          # 16 entries only
    tr.update(4)
    m = tr.metrics()
    assert m["episodes"] > 16 and m["episodes_dropped"] == m["episodes"] - 16 and len(m["episode_log"]) == 16
    assert abs(m["mean_return"] - m["mean_length"]) < 1e-9             # CartPole: return == length, from the exact sums
    big = drl.PPOTrainer(drl.PPOConfig(num_envs=256, num_steps=64, seed=2))
    big.update(4)
    mb = big.metrics()
    assert mb["episodes"] == m["episodes"] and mb["episodes_dropped"] == 0 and len(mb["episode_log"]) == mb["episodes"]


def _mann_whitney_p(x, y):
    """Two-sided Mann-Whitney U test, normal approximation with tie correction (n = 10 + 10 is large enough for it)."""
    from scipy import stats
    return float(stats.mannwhitneyu(x, y, alternative="two-sided").pvalue)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_learning_curve_distribution_matches_the_port(prec):
    """Learning-curve parity as a DISTRIBUTION (the RNG streams differ by contract): ten seeds of this build at the
    reference's shape (1 env x 128 steps x 156 updates, deep_rl/ppo.py:62-76) against ten seeds of the CPU port, which is
    bit-identical to the unmodified reference at the reference's seed (tests/golden/make_port_curves.py).  The last-20-episode
    mean returns must not be separable: Mann-Whitney p > 0.05 and |difference of means| < 1 pooled standard deviation."""
    import deep_rl_b200 as drl
    port = np.load(os.path.join(os.path.dirname(__file__), "golden", "port_curves_10seeds.npz"))
    want_last, want_first = port["last20"], port["first20"]
    last, first = [], []
    for seed in range(1, 11):
        cfg = drl.PPOConfig(seed=seed, update_precision=prec)
        tr = drl.PPOTrainer(cfg)
        eps = []
        for _ in range(cfg.num_updates()):
            tr.update()
            eps += [e[2] for e in tr.metrics()["episode_log"]]
        first.append(float(np.mean(eps[:20])))
        last.append(float(np.mean(eps[-20:])))
    last, first = np.array(last), np.array(first)
    pooled = float(np.sqrt(0.5 * (last.var(ddof=1) + want_last.var(ddof=1))))
    p = _mann_whitney_p(last, want_last)
    print(f"{prec}: last-20 means {np.round(last, 1)} (mean {last.mean():.1f}) vs port {np.round(want_last, 1)} (mean {want_last.mean():.1f}); "
          f"pooled sd {pooled:.1f}, Mann-Whitney p {p:.3f}")
    assert p > 0.05, (p, last, want_last)
    assert abs(last.mean() - want_last.mean()) < pooled, (last.mean(), want_last.mean(), pooled)
    assert abs(first.mean() - want_first.mean()) < 10.0            # untrained policies: ~22-28 steps per episode on both sides
    assert last.min() > 2.5 * first.mean()                         # every seed learns


@pytest.mark.parametrize("env_id,N,T", [("CartPole-v1", 256, 16), ("Acrobot-v1", 96, 24)])
def test_cuda_graph_update_is_bit_identical_to_eager(env_id, N, T):
    """The graph-replayed update (device-resident counters, drl_*_ctl entry points) must not change a single bit against the
    eager path with by-value counters: parameters, Adam moments, loss terms, env state, rollout buffers, episode log."""
    import deep_rl_b200 as drl
    out = []
    for graph in (False, True):
        cfg = drl.PPOConfig(env_id=env_id, num_envs=N, num_steps=T, seed=6, total_timesteps=N * T * 12, update_precision="bf16",
                            cuda_graph=graph)
        tr = drl.PPOTrainer(cfg)
        logs = []
        for u in range(7):
            tr.update(12)
            if u in (1, 4, 6):
                m = tr.metrics()
                logs.append((m["episodes"], m["loss"], m["grad_norm"], m["episode_log"]))
        torch.cuda.synchronize()
        assert (tr._graph is not None) == graph
        out.append((tr.agent.flat_params.clone(), tr.exp_avg.clone(), tr.exp_avg_sq.clone(), tr.loss_terms.clone(), tr.env.state.clone(),
                    tr.observations.clone(), tr.actions.clone(), tr.advantages.clone(), logs, tr.env.step_count, tr.adam_step,
                    tr.global_step))
    for a, b in zip(out[0], out[1]):
        if isinstance(a, torch.Tensor):
            assert torch.equal(a, b)
        else:
            assert a == b
    # a checkpoint taken from the graph-replaying trainer resumes bit-identically in an eager one
    g = drl.PPOTrainer(drl.PPOConfig(env_id=env_id, num_envs=N, num_steps=T, seed=6, total_timesteps=N * T * 12, update_precision="bf16"))
    for _ in range(4):
        g.update(12)
    sd = g.state_dict()
    g.update(12)
    e = drl.PPOTrainer(drl.PPOConfig(env_id=env_id, num_envs=N, num_steps=T, seed=6, total_timesteps=N * T * 12, update_precision="bf16",
                                     cuda_graph=False))
    e.load_state_dict(sd)
    e.update(12)
    torch.cuda.synchronize()
    assert torch.equal(g.agent.flat_params, e.agent.flat_params) and torch.equal(g.env.state, e.env.state)


def test_metrics_async_reads_what_metrics_reads():
    """Pipelined reads (update k's metrics consumed while update k+1 runs) return exactly what the synchronous metrics() returns."""
    import deep_rl_b200 as drl
    mk = lambda: drl.PPOTrainer(drl.PPOConfig(num_envs=512, num_steps=32, seed=8, total_timesteps=512 * 32 * 12, update_precision="bf16"))
    a, b = mk(), mk()
    sync = []
    for _ in range(6):
        a.update(12)
        m = a.metrics(with_episode_log="arrays")
        sync.append((m["loss"], m["grad_norm"], m["episodes"], m["mean_return"], sorted(zip(m["episode_log"]["step"].tolist(),
                     m["episode_log"]["env"].tolist(), m["episode_log"]["ret"].tolist(), m["episode_log"]["len"].tolist()))))
    def consume(h):      # the record arrays are views of a pinned buffer that the next-but-one read reuses: copy them out
        m = h.result()
        m["episode_log"] = {k: v.copy() for k, v in m["episode_log"].items()}
        return m

    got, pending = [], None
    for _ in range(6):
        b.update(12)
        h = b.metrics_async()
        if pending is not None:
            got.append(consume(pending))
        pending = h
    got.append(consume(pending))
    assert len(got) == 6
    for s, m in zip(sync, got):
        assert m["episodes_dropped"] == 0
        e = m["episode_log"]
        assert s == (m["loss"], m["grad_norm"], m["episodes"], m["mean_return"],
                     sorted(zip(e["step"].tolist(), e["env"].tolist(), e["ret"].tolist(), e["len"].tolist())))


@pytest.mark.parametrize("env_id,N,T", [("CartPole-v1", 300, 24), ("CartPole-v1", 128, 40), ("CartPole-v1", 1, 16), ("Acrobot-v1", 130, 16),
                                       ("MountainCar-v0", 96, 24)])
def test_rollout_h256_vs_oracle(env_id, N, T):
    """rollout256_kernel (actor, fused) + the batched critic pass: same teacher-forced checks as the 64-wide kernels."""
    _check_rollout(env_id, N, T, seed=3, tc=True, hidden=256)


def test_h256_trainer_learns_and_keeps_ratio_one():
    """Whole updates at hidden = 256 (BASELINE config C5's network): ratio = 1 at the first minibatch, finite losses, the
    policy improves; state_dict keys and shapes are the reference's with 64 -> 256 (deep_rl/ppo.py:34-47)."""
    import deep_rl_b200 as drl
    cfg = drl.PPOConfig(num_envs=1024, num_steps=64, hidden=256, seed=1, total_timesteps=1024 * 64 * 40)
    tr = drl.PPOTrainer(cfg)
    assert tr.update_precision == "bf16" and tr.rollout_precision == "bf16"
    sd = tr.agent.state_dict()
    assert sd["actor.2.weight"].shape == (256, 256) and sd["critic.4.weight"].shape == (1, 256) and sd["actor.0.weight"].shape == (256, 4)
    rets = []
    for u in range(40):
        tr.update()
        if u == 0:
            torch.cuda.synchronize()
            t0 = tr.loss_terms.cpu().numpy()
            assert abs(t0[0, 4]) < 1e-5 and t0[0, 5] == 0.0, t0[0]
        m = tr.metrics(with_episode_log=False)
        assert np.isfinite(m["loss"]) and np.isfinite(m["grad_norm"])
        if m["episodes"]:
            rets.append(m["mean_return"])
    assert rets[0] < 40 and max(rets[-5:]) > 100, rets
    with pytest.raises(ValueError, match="tensor-core"):
        drl.PPOTrainer(drl.PPOConfig(num_envs=8, hidden=256, update_precision="fp32"))


@pytest.mark.parametrize("env_id,N,T", [("CartPole-v1", 256, 32), ("Acrobot-v1", 64, 24)])
def test_epoch_per_launch_is_bit_identical_to_one_minibatch_per_launch(env_id, N, T):
    """drl_ppo_minibatch_update*(num_steps = 4): a whole epoch of ppo.py:156-192 inside one cooperative launch (weights reloaded and
    the record gather running ahead between the minibatches) must equal four single-minibatch launches bit for bit."""
    import deep_rl_b200 as drl
    out = []
    for spl in (1, 2, 4):
        cfg = drl.PPOConfig(env_id=env_id, num_envs=N, num_steps=T, seed=9, total_timesteps=N * T * 8, update_precision="bf16",
                            steps_per_launch=spl, cuda_graph=(spl == 4))
        tr = drl.PPOTrainer(cfg)
        assert tr.mb_per_launch == spl
        for _ in range(5):
            tr.update(8)
        torch.cuda.synchronize()
        out.append((tr.agent.flat_params.clone(), tr.exp_avg_sq.clone(), tr.loss_terms.clone(), tr.grad.clone(), tr.grad_norm.clone()))
    for o in out[1:]:
        for a, b in zip(out[0], o):
            assert torch.equal(a, b)
