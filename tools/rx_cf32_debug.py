"""Debug aid: the receive bank on cu8, the same samples widened to cf32, and the widened samples scaled down."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import __graft_entry__ as G
import bench

b200 = G.load_package()
b200.init(0)
dev = torch.device("cuda:0")
base = bench.c3_base_iq(0)
n_ch = 64
idx = np.arange(n_ch) % 16
taps = bench._p25_filter_taps()


def run(kind, scale=1.0, submit=False):
    rx = b200.P25p1Rx(n_ch, taps, rate_hz=48000, block_pairs=8192, max_pairs_per_call=49152, input_cu8=(kind == "cu8"), max_hits=32)
    out = rx.alloc_device_out(dev)
    good = tot = 0
    for t in range(5):
        tile = base[idx, t * 49152:(t + 1) * 49152]
        if kind == "cu8":
            d = torch.from_numpy(np.ascontiguousarray(tile)).to(dev)
        else:
            w = ((tile.astype(np.float32) - np.float32(127.5)) * np.float32(1.0 / 127.5)) * np.float32(scale)
            d = torch.from_numpy(np.ascontiguousarray(w)).to(dev)
        if submit:
            tk = rx.submit(d, 49152, out)
            rx.wait(tk)
        else:
            rx.process(d, 49152, out)
        fr, vo = rx.records(out)
        ok = ((fr["nid_status"] > 0) & (((fr["duid"] == 7) & (fr["n_tsbk"] > 0)) | ((fr["rs_kind"] > 0) & (fr["rs_status"] < 2)) | (fr["duid"] == 3))).sum()
        good += int(ok)
        tot += fr.size
    print(kind, scale, "submit" if submit else "process", "frames", tot, "ok", good, flush=True)


run("cu8")
run("cf32", 1.0)
run("cf32", 1.0, True)
run("cf32", 0.013)
run("cf32", 0.004)
