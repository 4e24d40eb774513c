"""Developer timing of the general channelizer (pfbn_kernel): python tools/pfbn_bench.py
Per configuration: n_out output times of an M-channel cu8 / cf32 wideband tile, bin stride R; CUDA events, 10 launches."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import json

import torch

import __graft_entry__ as G

b200 = G.load_package()
b200.init(0)
dev = torch.device("cuda:0")
rows = []
for M, R, cu8, n_out in [(256, 1, True, 49152), (1024, 1, True, 49152), (1024, 1, False, 49152), (2048, 2, True, 49152), (4096, 4, True, 24576),
                         (8192, 8, True, 24576), (4096, 1, True, 12288), (8192, 1, True, 6144), (512, 1, True, 49152), (2048, 1, True, 24576)]:
    n = n_out * M
    x = (torch.randint(0, 256, (n, 2), dtype=torch.uint8, device=dev) if cu8 else torch.randn((n, 2), device=dev))
    cz = b200.Channelizer(M, 8, cu8)
    out = torch.empty((M // R, n_out, 2), device=dev)
    for _ in range(3):
        cz.channelize_bins(x, R, R - 1, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        cz.channelize_bins(x, R, R - 1, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byts = n * (2 if cu8 else 8) + out.numel() * 4
    rows.append({"M": M, "R": R, "cu8": cu8, "n_out": n_out, "ms": round(ms, 4), "GS/s": round(n / ms / 1e6, 2), "GB/s": round(byts / ms / 1e6, 1)})
    print(json.dumps(rows[-1]), flush=True)
    del x, out, cz
