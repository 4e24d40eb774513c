"""Developer micro-benchmark (not the judged bench): times the block-side kernels with CUDA events."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

b200 = g.load_package()
b200.init(0)


def time_it(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return float(np.median(ts)), float(np.min(ts))


for n_ch, bp, nb in [(256, 8192, 6), (1024, 8192, 6), (4096, 8192, 2)]:
    for arith in (0, 1):
        # 4-level FSK (random symbols, 10 samples/symbol, +-1.8 kHz outer deviation at 48 kS/s) + noise, made on the GPU
        nsym = bp * nb // 10 + 1
        lv = torch.tensor([1.0, 3.0, -1.0, -3.0], device="cuda")[torch.randint(0, 4, (n_ch, nsym), device="cuda")]
        ph = torch.cumsum(lv.repeat_interleave(10, dim=1)[:, : bp * nb] * 0.0785, dim=1)
        iq = torch.stack([0.85 * torch.cos(ph), 0.85 * torch.sin(ph)], dim=-1) + 0.05 * torch.randn((n_ch, bp * nb, 2), device="cuda")
        iq = iq.contiguous().float()
        del ph, lv
        out = torch.empty((n_ch, bp * nb), device="cuda", dtype=torch.float32)
        bank = b200.DemodBank(n_ch, 48000, True, fir_arith=arith)
        med, mn = time_it(lambda: bank.full_demod(iq, bp, nb, out))
        samples = n_ch * bp * nb
        print(f"full_demod ch={n_ch} bp={bp} nb={nb} arith={'fma' if arith == 0 else 'nofma'}: "
              f"median {med:.3f} ms  min {mn:.3f} ms  -> {samples / med / 1e6:.2f} GS/s, "
              f"{samples * 12 / med / 1e6:.1f} GB/s algorithmic, x{samples / med * 1e3 / (n_ch * 48000):.0f} real time")
        bank.close()
        del iq, out

# channelizer: C2 shape, 256 channels x 49152 outputs (12.58 M wideband samples = 1.024 s at 12.288 MS/s)
for T in (8, 16):
    for cu8 in (False, True):
        n_out = 49152
        if cu8:
            x = torch.randint(0, 256, (n_out * 256, 2), device="cuda", dtype=torch.uint8)
        else:
            x = torch.randn((n_out * 256, 2), device="cuda", dtype=torch.float32)
        y = torch.empty((256, n_out, 2), device="cuda", dtype=torch.float32)
        cz = b200.Channelizer(256, T, cu8)
        med, mn = time_it(lambda: cz.channelize(x, y))
        n = n_out * 256
        bytes_alg = n * ((2 if cu8 else 8) + 8)
        print(f"channelize T={T} cu8={cu8}: median {med:.3f} ms min {mn:.3f} ms -> {n / med / 1e6:.2f} GS/s, "
              f"{bytes_alg / med / 1e6:.0f} GB/s algorithmic ({bytes_alg / mn / 1e6 / 6549.4 * 100:.1f}% of measured HBM peak)")
        cz.close()
        del x, y
