"""Debug aid: per-channel SNR of (synthesis bank -> cu8 -> channelizer) against the source channels. python tools/onestream_debug.py [M]"""
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
from dsdneo_b200 import shard

dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = 8
n = 16384
base = bench.c3_base_iq(0)[:, :n]
bc = torch.from_numpy((base.astype(np.float32) - 127.5) / 127.5).to(dev)
bc = torch.complex(bc[..., 0], bc[..., 1]).contiguous()
idx = (torch.arange(M, device=dev) * 7 + 3) % 16
cz = b200.Channelizer(M, T, True)
wide = shard.synthesize_wideband(torch, bc, idx, M, cz.prototype(), T)
print("wide u8 stats: min %d max %d clipped %.2e mean %.2f std %.2f" % (int(wide.min()), int(wide.max()),
      float(((wide == 0) | (wide == 255)).float().mean()), float(wide.float().mean()), float(wide.float().std())))
y = cz.channelize(wide)
y = torch.complex(y[..., 0], y[..., 1])  # [M, n]
res = []
for d in (T - 1,):
    ref = torch.roll(bc, d, dims=1)[idx]  # [M, n]
    sl = slice(64, n - 64)
    g = (ref[:, sl].conj() * y[:, sl]).sum(1) / (ref[:, sl].abs() ** 2).sum(1)
    err = y[:, sl] - g[:, None] * ref[:, sl]
    snr = 10 * torch.log10(((g.abs() ** 2)[:, None] * (ref[:, sl].abs() ** 2)).sum(1)) - 10 * torch.log10((err.abs() ** 2).sum(1))
    print("delay", d, "SNR dB: min %.1f median %.1f max %.1f; gain median %.5f" % (float(snr.min()), float(snr.median()), float(snr.max()), float(g.abs().median())))
    bad = torch.nonzero(snr < snr.median() - 6).flatten().tolist()
    print("channels 6 dB below the median:", bad[:40], len(bad))



del wide, y
n = 49152
NT = 5
base = bench.c3_base_iq(0)
bc = torch.from_numpy((base.astype(np.float32) - 127.5) / 127.5).to(dev)
bc = torch.complex(bc[..., 0], bc[..., 1]).contiguous()
wide = shard.synthesize_wideband(torch, torch.roll(bc, -(T - 1), dims=1).contiguous(), idx, M, cz.prototype(), T)
taps = bench._p25_filter_taps()
czz = b200.Channelizer(M, T, True)
rxs = {k: b200.P25p1Rx(M, taps, rate_hz=48000, block_pairs=8192, max_pairs_per_call=49152, input_cu8=False, max_hits=32) for k in "ACH"}
outs = {k: rxs[k].alloc_device_out(dev) for k in rxs}
for t in range(2 * NT):
    tt = t % NT
    yc = czz.channelize(wide[tt * n * M:(tt + 1) * n * M])
    sc = bc[:, tt * n:(tt + 1) * n][idx]
    s = (torch.stack([sc.real, sc.imag], dim=-1) * 0.0129).contiguous()
    h = yc.clone()
    if t == 0:
        h[:, :64] = s[:, :64]
    res = []
    for k, x in (("A", s), ("C", yc), ("H", h)):
        rxs[k].process(x, n, outs[k])
        f, _ = rxs[k].records(outs[k])
        res.append("%s %d/%d" % (k, int((f["nid_status"] > 0).sum()), f.size))
    print("tile", t, " ".join(res), flush=True)
