"""Stage timeline of the P25p1 receive bank's tile pipeline (debug aid): DSDNEO_B200_RX_TRACE=1 python tools/rx_trace.py
Prints, per traced tile, when each of the four stages started and ended (ms since the first traced tile entered stage A)."""
import ctypes as C
import os
import sys

os.environ["DSDNEO_B200_RX_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib

import numpy as np
import torch

import bench

b200 = importlib.import_module("dsd-neo_b200")


def main():
    dev = torch.device("cuda:0")
    base = bench.c3_base_iq(seed=0)
    idx = np.arange(bench.C3_CH) % base.shape[0]
    tiles = [torch.from_numpy(np.ascontiguousarray(base[idx, t * bench.C3_PAIRS:(t + 1) * bench.C3_PAIRS])).to(dev) for t in range(bench.C3_TILES)]
    rx = b200.P25p1Rx(bench.C3_CH, bench._p25_filter_taps(), rate_hz=bench.C3_RATE, block_pairs=bench.C3_BLOCK,
                      max_pairs_per_call=bench.C3_PAIRS, input_cu8=True, max_hits=32)
    host = len(sys.argv) > 1 and sys.argv[1] == "host"
    L = b200.lib()
    L.dsdneo_b200_p25p1_rx_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    n = 16
    if not host:
        out = rx.alloc_device_out(dev)
        last = -1
        for i in range(10):
            last = rx.submit(tiles[i % len(tiles)], bench.C3_PAIRS, out)
        rx.wait(last)
        torch.cuda.synchronize()
        L.dsdneo_b200_p25p1_rx_trace(rx._h, None, 0)
        for i in range(n):
            last = rx.submit(tiles[(10 + i) % len(tiles)], bench.C3_PAIRS, out)
        rx.wait(last)
    else:  # the host-buffer streaming form: pinned cu8 in, records + dibits out, five tiles in flight
        h_tiles = [t.cpu().pin_memory() for t in tiles]
        outs = [rx.alloc_host_out() for _ in range(6)]
        pending = []

        def run(i0, cnt):
            for i in range(i0, i0 + cnt):
                if len(pending) == 5:
                    rx.wait_host(pending.pop(0))
                pending.append(rx.submit_host(h_tiles[i % len(h_tiles)], bench.C3_PAIRS, outs[i % 6]))

        run(0, 10)
        while pending:
            rx.wait_host(pending.pop(0))
        torch.cuda.synchronize()
        L.dsdneo_b200_p25p1_rx_trace(rx._h, None, 0)
        run(10, n)
        while pending:
            rx.wait_host(pending.pop(0))
    ms = np.zeros((32, 10), np.float32)
    k = L.dsdneo_b200_p25p1_rx_trace(rx._h, ms.ctypes.data, 32)
    print("tile    A0     A1 |    B0     B1 |    C0     C1 |    D0     D1 |  H2D0   H2D1")
    for t in range(k):
        print("%4d %s" % (t, " | ".join("%6.2f %6.2f" % (ms[t, 2 * j], ms[t, 2 * j + 1]) for j in range(5))))
    print("steady-state ms per tile:", (ms[k - 1, 7] - ms[4, 7]) / (k - 1 - 4))


if __name__ == "__main__":
    main()
