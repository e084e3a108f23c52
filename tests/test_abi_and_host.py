"""CPU: the C-ABI library loads and exports every symbol include/drl_b200.h declares (no compute calls
without a GPU), argument validation returns error codes, and the host-side logic (config arithmetic,
env sharding, gloo world-2 gradient averaging, Philox world-size invariance) behaves."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _L():
    from deep_rl_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol():
    L = _L()
    header = open(os.path.join(ROOT, "include", "drl_b200.h")).read()
    declared = set(re.findall(r"\b(drl_[a-z_0-9]+)\s*\(", header))
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name), name
    nm = subprocess.run(["nm", "-D", "--defined-only", L.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (drl_[a-z_0-9]+)", nm))
    assert declared <= exported
    assert lib.drl_abi_version() == 2


def test_library_is_sm100a_only():
    L = _L()
    out = subprocess.run(["cuobjdump", "-lelf", L.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_shape_queries_and_error_codes():
    L = _L()
    lib = L.lib()
    assert (lib.drl_env_obs_dim(0), lib.drl_env_num_actions(0), lib.drl_env_obs_stride(0)) == (4, 2, 4)
    assert (lib.drl_env_obs_dim(1), lib.drl_env_num_actions(1), lib.drl_env_obs_stride(1)) == (6, 3, 8)
    assert (lib.drl_env_obs_dim(2), lib.drl_env_num_actions(2), lib.drl_env_obs_stride(2)) == (2, 3, 4)      # MountainCar-v0
    assert lib.drl_param_count(C.byref(L.NetT(2, 64, 3, 4))) == 8964
    cart, acro = L.NetT(4, 64, 2, 4), L.NetT(6, 64, 3, 8)
    assert lib.drl_param_count(C.byref(cart)) == 9155
    assert lib.drl_param_count(C.byref(acro)) == 9476
    assert lib.drl_record_width(C.byref(cart)) == 8 and lib.drl_record_width(C.byref(acro)) == 16
    assert lib.drl_workspace_bytes(C.byref(cart)) > 160 * 9155 * 4
    wide = L.NetT(4, 256, 2, 4)                                 # BASELINE config C5: 256-wide MLP
    assert lib.drl_param_count(C.byref(wide)) == 134_915
    assert lib.drl_workspace_bytes(C.byref(wide)) > (512 << 20)  # + the h1 / dz2 staging buffers of update256.cu
    bad = L.NetT(4, 128, 2, 4)
    assert lib.drl_param_count(C.byref(bad)) == -1 and b"hidden=128" in lib.drl_last_error()
    # argument validation happens before any CUDA call, so these are safe without a GPU
    assert lib.drl_env_reset(None, 0, 0) == -1
    env = L.EnvT(7, 4, 1, 0, 500, 0, 0, 0, 0)
    assert lib.drl_env_step(C.byref(env), 0, 0, 0, 0, 0, None, 0) == -1 and b"env kind" in lib.drl_last_error()
    assert lib.drl_pack_params(C.byref(bad), 0, 0, 0) == -2
    assert lib.drl_permutation(0, 16, 1, 0, 0, 0) == -1
    assert lib.drl_clip_adam(C.byref(cart), 0, 0, 0, 0, 1, 1e-3, 0.9, 0.999, 1e-5, 0.5, 1.0, 0, 0, 0) == -1
    with pytest.raises(L.DrlError):
        L.check(-1)


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import deep_rl_b200 as drl
    with pytest.raises(Exception, match="no CPU fallback"):
        drl.make("CartPole-v1")
    with pytest.raises(Exception, match="no CPU fallback"):
        drl.PPOTrainer(drl.PPOConfig())


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "deep_rl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "libdrl_oracle" not in src and "drl_or_" not in src, f


def test_config_arithmetic_matches_reference(golden):
    from deep_rl_b200 import PPOConfig
    c = PPOConfig()
    h = golden["hyper"]   # gamma, lambda, lr, clip, ent, vf, max_grad_norm, num_steps, num_updates, minibatch_size, epochs, seed
    assert (c.gamma, c.gae_lambda, c.learning_rate, c.clip_coef, c.ent_coef, c.vf_coef, c.max_grad_norm) == tuple(h[:7])
    assert (c.num_steps, c.num_updates(), c.minibatch_size, c.update_epochs, c.seed) == tuple(int(x) for x in h[7:])
    c = PPOConfig(num_envs=4096, total_timesteps=4096 * 128 * 10)
    assert c.batch_size == 524288 and c.minibatch_size == 131072 and c.num_updates() == 10 and c.num_updates(world=2) == 5


def test_shard_envs():
    from deep_rl_b200.dist import shard_envs
    assert [shard_envs(65536 * 8, r, 8) for r in (0, 7)] == [(0, 65536), (7 * 65536, 65536)]
    with pytest.raises(ValueError):
        shard_envs(10, 0, 4)


def test_oracle_world_size_invariance():
    """Env streams are keyed by the GLOBAL env id: two ranks of 4 envs == one rank of 8 envs."""
    from oracle import clib
    full = clib.OracleVecEnv("CartPole-v1", 8, seed=9)
    r0 = clib.OracleVecEnv("CartPole-v1", 4, seed=9, env_gid0=0)
    r1 = clib.OracleVecEnv("CartPole-v1", 4, seed=9, env_gid0=4)
    np.testing.assert_array_equal(full.reset(), np.concatenate([r0.reset(), r1.reset()]))
    rng = np.random.default_rng(0)
    for t in range(300):
        a = rng.integers(0, 2, size=8).astype(np.int32)
        of, rf, df, _ = full.step(a)
        o0, _, d0, _ = r0.step(a[:4])
        o1, _, d1, _ = r1.step(a[4:])
        np.testing.assert_array_equal(of, np.concatenate([o0, o1]))
        np.testing.assert_array_equal(df, np.concatenate([d0, d1]))
    lg = rng.normal(size=(8, 2)).astype(np.float32)
    np.testing.assert_array_equal(clib.sample(lg, 9, 0, 5)[0][4:], clib.sample(lg[4:], 9, 4, 5)[0])


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch
from deep_rl_b200 import dist
from oracle import ppo_oracle as po, clib
rank, world = dist.init_from_env(backend="gloo")
assert world == 2
rng = np.random.default_rng(0)
O, A, B = 4, 2, 64
params = po.init_params(O, 64, A, 1).numpy()
obs = rng.normal(size=(B, O)).astype(np.float32); act = rng.integers(0, A, size=B)
logp = -rng.uniform(0.4, 1.0, size=B).astype(np.float32); adv = rng.normal(size=B).astype(np.float32)
val = rng.normal(size=B).astype(np.float32); ret = adv + val
gid0, n = dist.shard_envs(B, rank, world)
sl = slice(gid0, gid0 + n)
# fixed (global) advantage statistics so that the mean over ranks equals the single-process gradient
mean, std = torch.tensor(float(adv.mean())), torch.tensor(float(adv.std(ddof=1)))
_, g_local = po.minibatch_loss_and_grad(params, obs[sl], act[sl], logp[sl], adv[sl], ret[sl], val[sl], O, 64, A, adv_mean=mean, adv_std=std)
g = torch.tensor(g_local)
dist.all_reduce_sum(g)
g_avg = (g / world).numpy()
_, g_full = po.minibatch_loss_and_grad(params, obs, act, logp, adv, ret, val, O, 64, A, adv_mean=mean, adv_std=std)
np.testing.assert_allclose(g_avg, g_full, rtol=1e-5, atol=1e-7)
p1, _, _, norm = po.clip_adam(params, g_avg, np.zeros_like(params), np.zeros_like(params), 1, 2.5e-4)
t = torch.tensor(p1); dist.all_reduce_sum(t)
np.testing.assert_allclose(t.numpy() / world, p1, rtol=0, atol=1e-7)     # ranks stay in lock-step
# global advantage statistics: rank-local (mean, std) of two minibatches merged over ranks == statistics of the union
xs = [rng.normal(loc=0.3 * k, scale=1.0 + k, size=(world, 40 + 8 * k)).astype(np.float32) for k in range(2)]
st = torch.tensor([[float(x[rank].mean()), float(x[rank].std(ddof=1))] for x in xs], dtype=torch.float32)
dist.merge_minibatch_stats(st, torch.tensor([x.shape[1] for x in xs]))
want = np.array([[x.reshape(-1).mean(), x.reshape(-1).std(ddof=1)] for x in xs])
np.testing.assert_allclose(st.numpy(), want, rtol=2e-6, atol=1e-7)
assert dist.all_reduce_max(float(rank), "cpu") == 1.0
dist.barrier(); dist.shutdown()
sys.stdout.write("rank-%d-ok\n" % rank); sys.stdout.flush()
"""


def test_gloo_world2_gradient_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    import socket
    with socket.socket() as sk:      # a free port, so parallel / repeated runs never collide
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "rank-0-ok" in res.stdout and "rank-1-ok" in res.stdout
