"""Times the 64-wide tensor-core minibatch gradient alone (drl_ppo_minibatch_grad: ppo_grad_tc_kernel WITHOUT the in-kernel tail +
grad_reduce_kernel) on synthetic records.  Usage: python profiles/tools/grad64_probe.py [M ...]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deep_rl_b200 import _lib as L  # noqa: E402


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [131072, 2097152]
    d = torch.device("cuda:0")
    lib = L.lib()
    net = L.NetT(4, 64, 2, 4)
    P = int(lib.drl_param_count(C.byref(net)))
    torch.manual_seed(0)
    params = torch.randn(P, device=d) * 0.1
    packed = torch.zeros(int(lib.drl_packed_count(C.byref(net))), dtype=torch.float32, device=d)
    L.check(lib.drl_pack_params(C.byref(net), params.data_ptr(), packed.data_ptr(), L.stream_ptr()))
    nb = int(lib.drl_workspace_bytes(C.byref(net)))
    ws = torch.zeros(nb, dtype=torch.uint8, device=d)
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=d)
    for M in sizes:
        B = 4 * M
        rec = torch.randn(B, 8, device=d)
        rec[:, 7] = torch.randint(0, 2, (B,), device=d).to(torch.int32).view(torch.float32)
        rec[:, 4] = -0.7
        idx = torch.randperm(B, device=d).to(torch.int32)
        stats = torch.tensor([0.0, 1.0], device=d)
        grad = torch.zeros(P, device=d)
        terms = torch.zeros(8, device=d)
        cf = L.PpoCoefT(0.2, 0.01, 0.5)
        call = lambda j: L.check(lib.drl_ppo_minibatch_grad(C.byref(net), packed.data_ptr(), rec.data_ptr(), idx.data_ptr(), j * M, M, stats.data_ptr(),
                                                            C.byref(cf), grad.data_ptr(), terms.data_ptr(), ws.data_ptr(), nb, 1, L.stream_ptr()))
        for j in range(4):
            call(j)
        tot = 0.0
        for rep in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for j in range(4):
                call(j)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1) / 4
        ms = tot / 5
        flops = 3 * 17_792 * M
        print(f"M={M}: {ms * 1e3:.1f} us per minibatch gradient (kernel + reduce), {flops / ms / 1e9:.1f} TFLOP/s ({flops / ms / 1e9 / 1406.7:.3f})")


if __name__ == "__main__":
    main()
