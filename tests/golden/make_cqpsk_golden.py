"""Generates tests/golden/cqpsk.npz from the UNMODIFIED reference full_demod() (CQPSK symbol output kind) compiled into
oracle/_ref/libdsdneo_ref.so.  Run in the dev container (needs /root/reference for `make -C oracle ref`):

    python tests/golden/make_cqpsk_golden.py

Inputs are float16-rounded so the fixture stays small; outputs are the reference's float32 symbols and per-block counts."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _harness as H  # noqa: E402

CASES = [
    # sps, rate, block_pairs, n_blocks, snr, cfo, squelch
    (5, 24000, 1200, 5, 18.0, 0.02, 0.0),
    (4, 24000, 800, 8, 20.0, -0.01, 0.0),
    (10, 48000, 1500, 3, 15.0, 0.005, 0.0),
    (5, 24000, 1000, 6, 20.0, 0.01, 1e-3),
]


def main():
    out = {"n_cases": np.int32(len(CASES))}
    for i, (sps, rate, bp, nb, snr, cfo, sq) in enumerate(CASES):
        rng = np.random.default_rng(0xC0 + i)
        x, _ = H.synth_cqpsk_iq(rng, bp * nb // sps + 1, sps=sps, snr_db=snr, cfo=cfo, timing=0.4)
        x = x[:bp * nb].astype(np.float16).astype(np.float32)
        if sq > 0:
            x[bp * 2:bp * 4] *= np.float32(1.0 / 1024)
        r = H.RefCqpsk(rate=rate, symrate=rate // sps, sps=sps, squelch=sq)
        sym, counts = r.run(x, bp, nb)
        out[f"cfg{i}"] = np.array([sps, rate, bp, nb], np.int32)
        out[f"squelch{i}"] = np.float32(sq)
        out[f"iq{i}"] = x.astype(np.float16)
        out[f"sym{i}"] = sym
        out[f"counts{i}"] = counts
    np.savez_compressed(os.path.join(HERE, "cqpsk.npz"), **out)
    print("wrote cqpsk.npz:", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
