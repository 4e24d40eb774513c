"""Developer benchmark for configuration C4 (BASELINE.json configs[3]): N DMR base-station channels (default 4096), 1.024 s per step:

  front (timed on a random cu8 tile: its cost does not depend on the content)
      one wideband cu8 stream of N channels x 48 kS/s -> polyphase channelizer (pfbn_kernel) -> full_demod (channel LPF + discriminator)
  sample side + FEC (synthetic DMR BS traffic at discriminator level: slot 1 = voice superframes A..F, slot 2 = data bursts)
      dmr matched filter + getSymbol + 4FSK slicer -> symbol stream (history on the device, steps joined) -> BS DATA / BS VOICE sync hunt
      -> data burst cutter -> BPTC(196,96), Golay(20,8) slot type
      -> voice burst cutter (6 bursts per superframe) -> AMBE+2 3600x2450 frame ECC (3 frames per burst)
      -> mbe synthesis of as many frames from synthetic parameters (parity unpinned; the parameter dequantiser is not built)

Prints one JSON line with per-kernel CUDA-event times (library timers) and what was decoded.  Not the judged bench.
Usage: python tools/c4_bench.py [n_channels] [steps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C

import __graft_entry__ as g
import _harness as H
from test_mbe_ecc import _o, ambe_encode

b200 = g.load_package()
b200.init(0)
dev = torch.device("cuda:0")
n_ch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
N_SAMP = 49152  # front: six full_demod blocks of 8192 pairs
N_SYM_SAMP = 49150  # sample side: whole symbols, so the replayed tile keeps the slicer's symbol phase
VOICE_SYNC = "131111333113313313113313"
DATA_SYNC = "313333111331131131331131"
L = _o()
amap = None


def voice_burst(rng, frames49, first, slot_bit):
    """One BS voice burst: CACH (TACT with the slot bit), three AMBE+2 frames through the interleave schedule, sync (burst A) or EMB."""
    global amap
    if amap is None:
        amap = b200.ambe_2450_dibit_map()
    cach = np.zeros(24, np.uint8)
    cach[:7] = H.hamming_7_4_encode_bruteforce((1, slot_bit, 0, 0))
    cach[7:] = rng.integers(0, 2, 17)
    tx = np.array([cach[H.DMR_CACH_INTERLEAVE[i]] for i in range(24)], np.uint8)
    dib = np.zeros(144, np.int64)
    dib[:12] = (tx[0::2] << 1) | tx[1::2]
    seg = [(0, 12, 36, 0), (1, 48, 18, 0), (1, 90, 18, 18), (2, 108, 36, 0)]
    frs = [ambe_encode(L, d) for d in frames49]
    for f, off, cnt, m0 in seg:
        for i in range(cnt):
            hr, hc, lr, lc = amap[m0 + i]
            dib[off + i] = (int(frs[f][hr, hc]) << 1) | int(frs[f][lr, lc])
    dib[66:90] = [int(c) for c in VOICE_SYNC] if first else rng.integers(0, 4, 24)
    return dib


def base_channel(rng, taps):
    """1.024 s of BS traffic: bursts alternate slot 1 (voice superframes) / slot 2 (data bursts)."""
    n_bursts = (N_SAMP // 10 - 42) // 144
    parts, sent_voice, sent_data = [rng.integers(0, 4, 40)], [], []
    vb = 0
    for b in range(n_bursts):
        if b % 2 == 0:
            frames = rng.integers(0, 2, (3, 49)).astype(np.uint8)
            parts.append(voice_burst(rng, frames, vb % 6 == 0, 0))
            sent_voice.append(frames)
            vb += 1
        else:
            payload = rng.integers(0, 2, 96).astype(np.uint8)
            burst, _ = H.dmr_build_bs_data_burst(rng, payload, 5, 3, tact4=(1, 1, 0, 0))
            parts.append(burst)
            sent_data.append(payload)
    dib = np.concatenate(parts)
    dib = np.concatenate([dib, rng.integers(0, 4, N_SAMP // 10 + 2 - dib.size)])
    return H.synth_dmr_disc(rng, dib, taps, 10000.0, 600.0)[:N_SYM_SAMP], sent_voice, sent_data


rng = np.random.default_rng(44)
from test_gpu_symbolizer import _taps

taps = _taps()
base = [base_channel(rng, taps[1]) for _ in range(8)]
x = torch.from_numpy(np.stack([base[c % 8][0] for c in range(n_ch)])).to(dev)
sy = b200.Symbolizer(n_ch, 48000, 4800, filters=taps)
sy.set_class([b200.sym_class_from_synctype(H.SYNC_DMR_BS_DATA_POS, H.SYNC_DMR_BS_DATA_POS)] * n_ch)
fs = b200.FrameSync(n_ch, [(DATA_SYNC, 10), (VOICE_SYNC, 12)])
KEEP, DELAY = 2048, 1536  # symbols of history per channel; the hunt trails the slicer by eleven bursts (10 * 144 + 54 <= 1536)
ss = b200.SymbolStream(n_ch, KEEP, sy.out_pitch(N_SYM_SAMP))
MAX_HITS, VOICE_HITS, VB = 40, 8, 11  # sync hits per channel and step: all / voice superframes; bursts cut per superframe
lib = b200.lib()
cz = b200.Channelizer(n_ch, 8, True) if n_ch >= 256 and (n_ch & (n_ch - 1)) == 0 else None
wide = torch.randint(0, 256, (N_SAMP * n_ch, 2), dtype=torch.uint8, device=dev) if cz else None
chan = torch.empty((n_ch, N_SAMP, 2), dtype=torch.float32, device=dev)
bank = b200.DemodBank(n_ch, 48000, True)
disc = torch.empty((n_ch, N_SAMP), dtype=torch.float32, device=dev)
ar = torch.arange(MAX_HITS, device=dev)[None, :]


def step(check=False):
    if cz is not None:
        cz.channelize(wide, chan)
    else:
        chan.normal_()
    bank.full_demod(chan, 8192, 6, disc)
    # the slicer writes behind the stream history kept on the device; the sync hunt runs DELAY symbols behind it, so bursts and
    # voice superframes that straddle steps are cut whole, once (dsdneo_b200_symbol_stream_*)
    view = ss.run(sy, x, N_SYM_SAMP)
    hits = torch.zeros((n_ch, MAX_HITS, 2), dtype=torch.int32, device=dev)
    n_hits = torch.zeros(n_ch, dtype=torch.int32, device=dev)
    b200.check(lib.dsdneo_b200_frame_sync_search_batch(fs._h, view.d_symbols + 4 * (KEEP - DELAY), view.pitch, view.d_new, hits.data_ptr(),
                                                       MAX_HITS, n_hits.data_ptr(), None))
    b200.check(lib.dsdneo_b200_sync_hits_rebase(hits.data_ptr(), n_hits.data_ptr(), n_ch, MAX_HITS, KEEP - DELAY, None))
    dh, dn = b200.sync_hits_select(hits, n_hits, 10, MAX_HITS)
    vh, vn = b200.sync_hits_select(hits, n_hits, 12, VOICE_HITS)
    k = n_ch * MAX_HITS
    cut = {name: torch.zeros((k, w) if w else (k,), dtype=torch.uint8, device=dev)
           for name, w in (("cach24", 24), ("info196", 196), ("rel98", 98), ("slot_type20", 20), ("valid", 0))}
    b200.check(lib.dsdneo_b200_dmr_burst_cut_batch(view.d_dibits, view.pitch, view.d_reliability, view.pitch, view.d_valid, dh.data_ptr(),
                                                   dn.data_ptr(), n_ch, MAX_HITS, 0, cut["cach24"].data_ptr(), cut["info196"].data_ptr(),
                                                   cut["rel98"].data_ptr(), cut["slot_type20"].data_ptr(), cut["valid"].data_ptr(), None))
    out96 = torch.zeros((k, 96), dtype=torch.uint8, device=dev)
    r3 = torch.zeros((k, 3), dtype=torch.uint8, device=dev)
    errs = torch.zeros(k, dtype=torch.int32, device=dev)
    b200.check(lib.dsdneo_b200_bptc_196x96_batch(cut["info196"].data_ptr(), 1, out96.data_ptr(), r3.data_ptr(), errs.data_ptr(), k, None))
    slot_ok = torch.zeros(k, dtype=torch.uint8, device=dev)
    slot8 = torch.zeros((k, 8), dtype=torch.uint8, device=dev)
    b200.check(lib.dsdneo_b200_fec_block_decode_batch(b200.FEC_GOLAY_20_8, cut["slot_type20"].data_ptr(), slot8.data_ptr(), slot_ok.data_ptr(), k, None))
    R = n_ch * VOICE_HITS * VB
    cach = torch.zeros((R, 24), dtype=torch.uint8, device=dev)
    fr = torch.zeros((R, 3, 4, 24), dtype=torch.uint8, device=dev)
    sync = torch.zeros((R, 48), dtype=torch.uint8, device=dev)
    valid = torch.zeros(R, dtype=torch.uint8, device=dev)
    b200.check(lib.dsdneo_b200_dmr_voice_cut_batch(view.d_dibits, view.pitch, view.d_valid, vh.data_ptr(), vn.data_ptr(), n_ch, VOICE_HITS, VB, 0,
                                                   cach.data_ptr(), fr.data_ptr(), sync.data_ptr(), valid.data_ptr(), None))
    n_fr = fr.shape[0] * 3
    ambe_d = torch.zeros((n_fr, 49), dtype=torch.uint8, device=dev)
    c0 = torch.zeros(n_fr, dtype=torch.int32, device=dev)
    tot = torch.zeros(n_fr, dtype=torch.int32, device=dev)
    b200.check(lib.dsdneo_b200_ambe3600x2450_decode_batch(fr.data_ptr(), ambe_d.data_ptr(), c0.data_ptr(), tot.data_ptr(), n_fr, None))
    if not check:
        return None
    torch.cuda.synchronize()
    v = valid.view(n_ch, VOICE_HITS, VB).bool() & (torch.arange(VB, device=dev) % 2 == 0) & (ar[:, :VOICE_HITS] < vn[:, None])[..., None]
    vt = tot.view(n_ch, VOICE_HITS, VB, 3)
    data_ok = ((cut["valid"] == 1) & (errs == 0) & (slot_ok == 1))
    # channel 0 carries base[0]: every decoded payload / parameter vector must be one that was sent
    sent_p = {p.tobytes() for p in base[0][2]}
    got_p = out96[:MAX_HITS][data_ok[:MAX_HITS]].cpu().numpy()
    sent_f = {f.tobytes() for fr3 in base[0][1] for f in fr3}
    got_f = ambe_d.view(n_ch, VOICE_HITS, VB, 3, 49)[0][v[0]].cpu().numpy().reshape(-1, 49)
    return {"data_bursts_ok": int(data_ok.sum()), "data_bursts_valid": int((cut["valid"] == 1).sum()),
            "voice_bursts": int(v.sum()), "ambe_frames": int(v.sum()) * 3, "ambe_frames_clean": int((vt[v] == 0).sum()),
            "channel0_payloads_sent": len(sent_p), "channel0_payloads_recovered": sum(1 for p in got_p if p.tobytes() in sent_p),
            "channel0_ambe_sent": len(sent_f), "channel0_ambe_recovered": sum(1 for f in got_f if f.tobytes() in sent_f)}


for _ in range(3):
    step()
detail = step(check=True)
b200.timing_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
kt = b200.timing_report()
b200.timing_enable(False)
# MBE synthesis of as many frames as one step decodes (synthetic parameters; host entry point, its own timing)
n_syn = min(detail["ambe_frames"], 20000)
cur, prev = (b200.MbeParms * n_syn)(), (b200.MbeParms * n_syn)()
for i in range(n_syn):
    for p in (cur[i], prev[i]):
        p.w0, p.L = 0.04 + 0.0001 * (i % 50), 30
        for l in range(1, 31):
            p.Vl[l], p.Ml[l] = (l + i) % 3 != 0, 500.0 / l
b200.timing_enable(True)
b200.mbe_synth(cur, prev, want_int16=True)
ks = b200.timing_report()
b200.timing_enable(False)
print(json.dumps({
    "workload": "C4 developer line: %d DMR BS channels, 1.024 s per step" % n_ch, "ms_per_step": ms,
    "x_realtime": 1024.0 / ms, "channels_at_realtime": n_ch * 1024.0 / ms, "iq_msps": n_ch * N_SAMP / ms / 1e3,
    "kernels_ms": {k: v["ms"] / max(1, v["launches"]) * (v["launches"] / steps) for k, v in kt.items()},
    "mbe_synth": {"frames": n_syn, "kernels_ms": {k: v["ms"] for k, v in ks.items()}}, "decoded": detail}))
