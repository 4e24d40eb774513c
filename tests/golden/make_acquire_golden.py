"""Generates tests/golden/acquire.npz: the UNMODIFIED getFrameSync() (src/dsp/dsd_frame_sync.c, compiled into oracle/_ref with
the hook seam of oracle/ref_shim_symbol.c) on the reference's own IQ-replay captures, from the never-synchronised state, followed
by 600 getDibitSoft() calls.  Inputs: the cu8 captures (the two P25 ones are already held by c1_p25p1_c4fm_{cc,vc}.npz; the DMR
ones, tests/fixtures/iq/dmr_t3_cc.iq and dmr_voice.iq -- 2 s at 48 kS/s each, which the reference's CLI tests decode with the
DMR preset (rf_mod 2, 12.5 kHz channel filter) -- are stored here).  Run in the dev container:

    python tests/golden/make_acquire_golden.py

Stored per case: sync type, symbols hunted (saturating at the reference's 2048-entry history), samples consumed, the slicer
state right after the sync (min, max, center, umid, lmid, minref, maxref, lastsample), the newest <= 200 hunt symbols with their
rolling payload dibits / reliabilities (after resample-on-sync), and the first 600 synchronised dibits / reliabilities / LLRs /
symbols; for the -xr cases one record of the UNMODIFIED dmr_data_sync() per burst of the stream (what it hands
to dmr_data_burst_handler / dmr_cach, the slot-type colour code and the colour code its confidence gate locks)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _harness as H  # noqa: E402
from test_oracle_symbol import _ref_acquire  # noqa: E402

# name -> (capture, frame mask (1 P25p1, 2 DMR), rf_mod, channel LPF profile)
CASES = {"p25p1_c4fm_cc": ("p25p1_c4fm_cc", 1, 0, 4), "p25p1_c4fm_vc": ("p25p1_c4fm_vc", 1, 0, 4),
         "dmr_t3_cc": ("dmr_t3_cc", 2, 2, 2), "dmr_voice": ("dmr_voice", 2, 2, 2), "dmr_t3_cc_c4fm": ("dmr_t3_cc", 3, 0, 2),
         "dmr_t3_cc_xr": ("dmr_t3_cc", 6, 2, 2), "dmr_voice_xr": ("dmr_voice", 6, 2, 2)}  # mask 4 = opts->inverted_dmr (-xr)
BP = 8000


def main():
    out = {}
    for name, (cap, mask, rf_mod, profile) in CASES.items():
        u = np.fromfile("/root/reference/tests/fixtures/iq/%s.iq" % cap, dtype=np.uint8)
        x = H.widen_cu8(u).reshape(-1, 2)
        disc = H.RefDemod("par", rate=48000, symrate=4800, profile=profile).run(x, BP, x.shape[0] // BP)
        ref = _ref_acquire(disc, mask, rf_mod)
        assert ref["sync_type"] >= 0
        if cap.startswith("dmr"):
            out["iq_" + cap] = u
        out[name + "_cfg"] = np.array([mask, rf_mod, profile, ref["sync_type"], ref["hunted"], ref["consumed"]], np.int32)
        out[name + "_f8"] = ref["f8"]
        out[name + "_recent_sym"] = ref["recent_sym"]
        out[name + "_recent_dib"] = ref["recent_dib"].astype(np.uint8)
        out[name + "_recent_rel"] = ref["recent_rel"]
        d, r, l, s = ref["after"]
        out[name + "_after_dib"], out[name + "_after_rel"], out[name + "_after_llr"], out[name + "_after_sym"] = d, r, l, s
        print(name, "sync", ref["sync_type"], "hunted", ref["hunted"], "consumed", ref["consumed"], "after", d.size)
        if name.endswith("_xr"):
            # the UNMODIFIED dmr_data_sync() (oracle/ref_shim_dmr.c) on every burst of the whole (pinned) oracle stream
            import test_acquire as T
            from test_frame_sync import oracle_search
            o = T._oracle_acquire(disc, mask, rf_mod, disc.size // 9)
            dib = np.concatenate([o["dib"], o["after"][0]])
            rel = np.concatenate([o["rel"], o["after"][1]])
            sym = np.concatenate([o["sym"], o["after"][3]])
            n, pos, _, _, _ = oracle_search(sym, [(H.DMR_BS_VOICE_STR, 13)], max_hits=80)
            pos = [int(p) for p in pos[:n] if p >= 89 and p + 55 <= dib.size]
            recs = H.ref_dmr_bursts(dib, rel, pos, True)
            out[name + "_burst_pos"] = np.array(pos, np.int32)
            out[name + "_burst_ref"] = recs.view(np.uint8).reshape(len(pos), -1)
            print("   bursts", len(pos), "handled", int(recs["handler_called"].sum()), "colour codes", sorted(set(recs["dmr_color_code"].tolist())))
    path = os.path.join(HERE, "acquire.npz")
    np.savez_compressed(path, **out)
    print(os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
