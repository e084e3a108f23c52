"""GPU: the tcgen05.mma operand layouts / descriptors used by the tensor-core update path, checked against
fp32 matmuls of the bf16-rounded operands (tolerance: fp32 accumulation order only)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(mode, variant, a, b, shape):
    from deep_rl_b200 import _lib as L
    d = torch.device("cuda:0")
    ta, tb = torch.tensor(a, device=d), torch.tensor(b, device=d)
    out = torch.full(shape, float("nan"), dtype=torch.float32, device=d)
    L.check(L.lib().drl_selftest_umma(mode, variant, ta.data_ptr(), tb.data_ptr(), out.data_ptr(), L.stream_ptr()))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _bf16(x):
    return torch.tensor(x).to(torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_umma_layouts(mode):
    rng = np.random.default_rng(mode)
    if mode == 0:
        a, b = rng.normal(size=(128, 64)), rng.normal(size=(64, 64))
        want = _bf16(a.astype(np.float32)) @ _bf16(b.astype(np.float32)).T
    elif mode == 1:
        a, b = rng.normal(size=(128, 64)), rng.normal(size=(64, 64))
        want = _bf16(a.astype(np.float32)) @ _bf16(b.astype(np.float32))
    elif mode == 2:
        a, b = rng.normal(size=(128, 128)), rng.normal(size=(128, 128))
        want = _bf16(a.astype(np.float32)).T @ _bf16(b.astype(np.float32))
    else:
        a, b = rng.normal(size=(128, 128)), rng.normal(size=(128, 16))
        want = _bf16(a.astype(np.float32)).T @ _bf16(b.astype(np.float32))
    a, b = a.astype(np.float32), b.astype(np.float32)
    errs = {}
    for variant in (0, 1):
        got = _run(mode, variant, a, b, want.shape)
        errs[variant] = float(np.nanmax(np.abs(got - want))) if np.isfinite(got).any() else float("inf")
        print(f"mode {mode} variant {variant}: max abs err {errs[variant]:.3e}, nan count {int(np.isnan(got).sum())}")
        if variant == 0 and errs[0] > 1e-3:
            # diagnosis aid: which output rows / cols are right
            ok = np.abs(got - want) < 1e-3
            print("  rows fully ok:", int(ok.all(1).sum()), "cols fully ok:", int(ok.all(0).sum()), "elements ok:", int(ok.sum()), "/", ok.size)
    assert errs[0] < 1e-3, errs
