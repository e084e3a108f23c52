"""Times the two kernels of the 256-wide minibatch gradient separately (CUDA events around drl_ppo_minibatch_grad calls on
synthetic records), for a sweep of minibatch sizes.  Usage: python profiles/tools/h256_probe.py [M ...]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deep_rl_b200 import _lib as L  # noqa: E402


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [262144, 1048576]
    d = torch.device("cuda:0")
    lib = L.lib()
    net = L.NetT(4, 256, 2, 4)
    P = int(lib.drl_param_count(C.byref(net)))
    torch.manual_seed(0)
    params = torch.randn(P, device=d) * 0.05
    packed = torch.zeros(int(lib.drl_packed_count(C.byref(net))), dtype=torch.float32, device=d)
    L.check(lib.drl_pack_params(C.byref(net), params.data_ptr(), packed.data_ptr(), L.stream_ptr()))
    nb = int(lib.drl_workspace_bytes(C.byref(net)))
    ws = torch.zeros(nb, dtype=torch.uint8, device=d)
    for M in sizes:
        rec = torch.randn(M, 8, device=d)
        rec[:, 7] = torch.randint(0, 2, (M,), device=d).to(torch.int32).view(torch.float32)
        rec[:, 4] = -0.7
        idx = torch.randperm(M, device=d).to(torch.int32)
        stats = torch.tensor([0.0, 1.0], device=d)
        grad = torch.zeros(P, device=d)
        terms = torch.zeros(8, device=d)
        cf = L.PpoCoefT(0.2, 0.01, 0.5)
        call = lambda: L.check(lib.drl_ppo_minibatch_grad(C.byref(net), packed.data_ptr(), rec.data_ptr(), idx.data_ptr(), 0, M, stats.data_ptr(),
                                                          C.byref(cf), grad.data_ptr(), terms.data_ptr(), ws.data_ptr(), nb, 1, L.stream_ptr()))
        for _ in range(2):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        flops = 3 * 267_776 * M
        print(f"M={M}: {ms:.3f} ms per minibatch gradient, {flops / ms / 1e9:.1f} TFLOP/s algorithmic ({flops / ms / 1e9 / 1406.7:.3f} of 1406.7)")


if __name__ == "__main__":
    main()
