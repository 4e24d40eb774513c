"""Developer probe: where the host-buffer path of the P25 receiver bank spends its step (H2D alone, device step alone, the
pipelined submit/wait loop with and without the dibit D2H)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
import bench as B

b200 = g.load_package()
b200.init(0)
base = B.c3_base_iq()
idx = np.arange(B.C3_CH) % base.shape[0]
h_tiles = []
for t in range(B.C3_TILES):
    ht = torch.empty((B.C3_CH, B.C3_PAIRS, 2), dtype=torch.uint8).pin_memory()
    ht.copy_(torch.from_numpy(np.ascontiguousarray(base[idx, t * B.C3_PAIRS:(t + 1) * B.C3_PAIRS])))
    h_tiles.append(ht)
d = torch.empty_like(h_tiles[0], device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20):
    d.copy_(h_tiles[i % 5], non_blocking=True)
torch.cuda.synchronize()
print("H2D alone ms", (time.perf_counter() - t0) / 20 * 1e3)
taps = B._p25_filter_taps()
for with_dibits in (True, False):
    rx = b200.P25p1Rx(B.C3_CH, taps, block_pairs=B.C3_BLOCK, max_pairs_per_call=B.C3_PAIRS)
    outs = [rx.alloc_host_out(with_dibits), rx.alloc_host_out(with_dibits)]

    def run(n):
        prev = None
        for i in range(n):
            t = rx.submit_host(h_tiles[i % 5], B.C3_PAIRS, outs[i % 2])
            if prev is not None:
                rx.wait_host(prev)
            prev = t
        rx.wait_host(prev)

    run(6)
    t0 = time.perf_counter()
    run(20)
    print("pipelined, dibits" if with_dibits else "pipelined, no dibits", (time.perf_counter() - t0) / 20 * 1e3)
    t0 = time.perf_counter()
    for i in range(10):
        rx.wait_host(rx.submit_host(h_tiles[i % 5], B.C3_PAIRS, outs[0]))
    print("  blocking", (time.perf_counter() - t0) / 10 * 1e3)

rx = b200.P25p1Rx(B.C3_CH, taps, block_pairs=B.C3_BLOCK, max_pairs_per_call=B.C3_PAIRS)
out = rx.alloc_device_out("cuda")
d_tiles = [t.cuda() for t in h_tiles]
for i in range(6):
    rx.process(d_tiles[i % 5], B.C3_PAIRS, out)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20):
    rx.process(d_tiles[i % 5], B.C3_PAIRS, out)
torch.cuda.synchronize()
print("device path, async ms", (time.perf_counter() - t0) / 20 * 1e3)
t0 = time.perf_counter()
for i in range(20):
    rx.process(d_tiles[i % 5], B.C3_PAIRS, out)
    torch.cuda.synchronize()
print("device path, sync each ms", (time.perf_counter() - t0) / 20 * 1e3)
b200.timing_enable(True)
for i in range(5):
    rx.process(d_tiles[i % 5], B.C3_PAIRS, out)
torch.cuda.synchronize()
rep = b200.timing_report()
print({k: round(v["ms"] / v["launches"], 3) for k, v in rep.items()})
