"""Generates tests/golden/c3_p25_dibits.npz: the transmitted dibit streams of the synthetic C3 workload of bench.py
(BASELINE.json configs[2]: 1024 synthetic P25 Phase 1 channels).  16 base channels x 24576 symbols (5.12 s at 4800 sym/s, five
bench tiles of 49152 samples, so the rotation closes on a whole symbol): 8 voice channels (HDU, 13 x {LDU1, LDU2} with random
IMBE payloads and valid Hamming / RS / LSD coding, TDU) and 8 control channels (68 three-block TSDUs with valid trellis
coding and CRC-16, last-block flag on the third block), status symbols inserted, every frame carrying a valid NID.  bench.py modulates them (C4FM shaping, FM,
AWGN, cu8) itself; only the frame ENCODERS of the test harness are needed here, which is why this is a fixture.

    python tests/golden/make_c3_dibits.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _harness as H  # noqa: E402

L = 24576


def tdu(nac):
    body = [np.array(H.P25P1_SYNC_DIBITS), H.p25p1_nid_dibits(nac, 0x3), np.zeros(14, np.int64)]
    return H.p25p1_insert_status(np.concatenate(body))


def main():
    rng = np.random.default_rng(0xC3)
    chans, kinds = [], []
    for c in range(16):
        nac = int(rng.integers(1, 0xFFE))
        parts = []
        if c < 8:
            parts.append(H.p25p1_build_hdu(rng, nac)[0])
            for _ in range(13):
                parts.append(H.p25p1_build_ldu(rng, nac, False)[0])
                parts.append(H.p25p1_build_ldu(rng, nac, True)[0])
            parts.append(tdu(nac))
        else:
            for _ in range(68):
                parts.append(H.p25p1_build_tsdu(rng, nac, 3, H._bch_nid_encoder(), valid_crc=True)[0])
        s = np.concatenate(parts)
        assert s.size <= L, s.size
        s = np.concatenate([s, rng.integers(0, 4, L - s.size)])  # idle tail
        chans.append(s.astype(np.uint8))
        kinds.append(0 if c < 8 else 1)
    out = os.path.join(HERE, "c3_p25_dibits.npz")
    np.savez_compressed(out, dibits=np.stack(chans), kind=np.array(kinds, np.uint8), symbols=np.int32(L))
    print("wrote", out, os.path.getsize(out), [int(np.sum(c < 4)) for c in chans][:2])


if __name__ == "__main__":
    main()
