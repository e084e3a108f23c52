"""Per-phase cycle stamps of mlp256_kernel (CTA 0, compute warp 0): DRL_TC_DEBUG=1 python profiles/tools/h256_stamps.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["DRL_TC_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deep_rl_b200 import _lib as L  # noqa: E402

NAMES = ["tile start", "L1 done", "P0 done", "fwd-a done", "P1a tanh done", "fwd-b done", "P1b + head partials", "quad barrier",
         "loss, h2 -> ring, W4RDY", "dz2 regs", "dW4 done", "dz2 stored, h1 reloaded", "dh1+db2 done", "dz1 stored (tile end)"]


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
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
    rec = torch.randn(M, 8, device=d)
    rec[:, 7] = torch.randint(0, 2, (M,), device=d).to(torch.int32).view(torch.float32)
    idx = torch.randperm(M, device=d).to(torch.int32)
    stats = torch.tensor([0.0, 1.0], device=d)
    grad = torch.zeros(P, device=d)
    terms = torch.zeros(8, device=d)
    cf = L.PpoCoefT(0.2, 0.01, 0.5)
    for _ in range(2):
        L.check(lib.drl_ppo_minibatch_grad(C.byref(net), packed.data_ptr(), rec.data_ptr(), idx.data_ptr(), 0, M, stats.data_ptr(),
                                           C.byref(cf), grad.data_ptr(), terms.data_ptr(), ws.data_ptr(), nb, 1, L.stream_ptr()))
    torch.cuda.synchronize()
    # debug block = 4 KB (1024-aligned) right before the staging region
    raw = ws.cpu().numpy()
    P4 = (P + 3) // 4 * 4
    off_debug = (64 + 8 * 16 * 512 * 2 + 8 * 1024 + 8 * 512 * 4 + 4 * 160 * 8 + 4 * 160 * P4 + 1023) // 1024 * 1024
    st = raw[off_debug:off_debug + 4096].view(np.int64)
    for k in range(2, 6):
        row = st[k * 16:k * 16 + 14]
        base = row[0]
        print(f"tile {k}: " + ", ".join(f"{NAMES[i]} +{int(row[i] - base)}" for i in range(1, 14)))
        if k + 1 < 8:
            print(f"    next tile starts +{int(st[(k + 1) * 16] - base)}")


if __name__ == "__main__":
    main()
