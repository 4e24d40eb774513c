"""Developer benchmark for the CQPSK symbol output kind: N synthetic P25 LSM-like channels at
24 kS/s (sps 5), 1 s of signal per step, channel LPF + AGC/FLL/Gardner/Costas chain + symbol-rate slicer.  Prints per-kernel CUDA-event times
(library timers) and derived rates; not the judged bench (that is bench.py = C2).
usage: python tools/cqpsk_bench.py [n_channels] [rate] [sps]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import _harness as H

b200 = g.load_package()
b200.init(0)
n_ch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rate = int(sys.argv[2]) if len(sys.argv) > 2 else 24000
sps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
bp, nb = rate // 10, 10  # ten 100 ms blocks = 1 s of signal
rng = np.random.default_rng(5)
base = [H.synth_cqpsk_iq(rng, bp * nb // sps + 2, sps=sps, snr_db=18.0, cfo=0.004 * (c - 4), timing=0.11 * c)[0][:bp * nb]
        for c in range(8)]
x = torch.from_numpy(np.stack([base[c % 8] for c in range(n_ch)])).cuda()
bank = b200.CqpskBank(n_ch, rate, ted_sps=[sps] * n_ch)
slicer = b200.CqpskSlicer(n_ch)


def step():
    sym, counts = bank.full_demod(x, bp, nb)
    res = slicer.run(sym, counts.sum(dim=1, dtype=torch.int32).contiguous())
    return sym, counts, res


for _ in range(3):
    sym, counts, res = step()
torch.cuda.synchronize()
b200.timing_enable(True)
iters = 10
t0 = time.perf_counter()
for _ in range(iters):
    sym, counts, res = step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / iters
rep = b200.timing_report()
b200.timing_enable(False)
n_sym = int(counts.sum().item())
n_samp = n_ch * bp * nb
out = {"channels": n_ch, "rate": rate, "sps": sps, "samples_per_channel": bp * nb, "symbols": n_sym,
       "wall_ms_per_step": wall * 1e3, "x_realtime": (bp * nb / rate) / wall,
       "kernels": {k: v["ms"] / v["launches"] for k, v in rep.items()}}
# algorithmic bytes: LPF reads cf32, writes cf32 (16 B/sample); chain reads cf32, writes one f32 per symbol
alg = {"lpf_phase_kernel": n_samp * 16.0, "cqpsk_chain_kernel": n_samp * 8.0 + n_sym * 4.0, "cqpsk_slice_kernel": n_sym * 10.0}
out["achieved_gbs"] = {k: alg[k] / (out["kernels"][k] * 1e-3) / 1e9 for k in alg if k in out["kernels"]}
out["ns_per_channel_sample"] = {k: out["kernels"][k] * 1e6 / (bp * nb) for k in alg if k in out["kernels"]}
print(json.dumps(out))
