#!/usr/bin/env python
"""bench.py -- the judged benchmark (contract in the task statement, section 4).

Default workload (BASELINE.json configs[2], "C3", the configuration the metric is quoted on): 1024 synthetic P25 Phase 1 C4FM
channels per GPU, end to end.  One step = 1.024 s of signal on every channel (49152 cu8 IQ pairs at 48 kS/s = 100.7 MB per
GPU, the reference's RTL ingest / --iq-replay sample format) through ONE C-ABI object, dsdneo_b200_p25p1_rx_*:
    widen_u8_to_f32_bias127 -> full_demod (channel LPF + FSK discriminator) -> p25_filter + getDibitSoft (symbol timing,
    threshold tracker, 4-level slicer, soft metrics) -> frame sync -> NID (BCH + Chase) -> processTSBK (half-rate trellis,
    list-8 + CRC) / processHDU (Golay(24,6), RS(36,20,17)) / processLDU1,2 (IMBE de-interleave, Hamming(10,6,3),
    RS(24,12,13) / RS(24,16,9), LSD)
Outputs: frame records, the de-interleaved IMBE frames the reference hands to mbelib-neo, and the dibit stream.  (The
vocoder itself is an un-vendored dependency of the reference: its parity cannot be pinned here, so it is not inside the
judged step; `--workload mbe` times the synthesis stage separately.)
The metric is IQ MS/s (complex channel samples consumed per second, whole job) and `channels_at_realtime`.

  python bench.py [--gpus N] [--steps K] [--warmup W]            this framework (CUDA, through the C-ABI)
  python bench.py --impl reference [...]                         the reference's own CPU code on the host cores, same workload
  python bench.py --workload c2 | cqpsk | fec [...]              developer lines (C2 = 256-channel channelizer + discriminator)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M = 256
WIDEBAND_HZ = 12_288_000
CHAN_HZ = WIDEBAND_HZ // M  # 48000
BLOCK_PAIRS = 8192  # the reference's DEFAULT_BUF_LENGTH (16384 floats) per full_demod() call
N_BLOCKS = 6
N_OUT = BLOCK_PAIRS * N_BLOCKS  # 49152 channel samples = 1.024 s
N_IN = N_OUT * M  # 12,582,912 wideband samples per step
SYMBOL_HZ = 4800
N_ROTATE = 3  # distinct input buffers: 3 x 100.7 MB > 126 MB L2
WORKLOAD = ("C2: 256-channel polyphase FIR channelizer + FM discriminator (channel LPF + FSK discriminator = "
            "reference full_demod) on synthetic 12.288 MS/s complex IQ, 1xB200 per 256-channel band")


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception:
            return self
        names = {
            "hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(pynvml, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }

        def run():
            while not self._stop.is_set():
                try:
                    self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                    try:
                        mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for k, bit in names.items():
                        if mask & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
                time.sleep(self.period)

        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()
        return self

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------------- data

def make_wideband(torch, device, seed: int):
    """Synthetic wideband band: 256 4-level FSK carriers (one per channelizer bin, 4800 sym/s, +-1.8 kHz outer
    deviation, independent random dibits seeded per channel) + AWGN at 30 dB per-carrier SNR, 0.5 full scale.
    Built on the GPU in float64 phase, stored cf32.  Returns [N_IN, 2] float32."""
    g = torch.Generator(device=device)
    g.manual_seed(0xD5D + seed)
    sps = WIDEBAND_HZ // SYMBOL_HZ  # 2560 wideband samples per symbol
    nsym = N_IN // sps + 1
    t = torch.arange(N_IN, device=device, dtype=torch.int64)
    acc = torch.zeros((N_IN, 2), device=device, dtype=torch.float32)
    levels = torch.tensor([1, 3, -1, -3], device=device, dtype=torch.int64)
    two_pi = 2.0 * torch.pi
    # phase is accumulated EXACTLY in integers: one level step = 600 Hz = 600/12.288e6 cycles per sample, and the
    # carrier k/M cycles per sample; 600/12288000 = 1/20480 -> work in units of 1/20480 cycle (M=256 divides 20480)
    unit = WIDEBAND_HZ // 600  # 20480 phase units per cycle
    for k in range(M):
        dib = torch.randint(0, 4, (nsym,), device=device, generator=g)
        lv = levels[dib].repeat_interleave(sps)[:N_IN]
        ph_units = torch.cumsum(lv, 0) + (t * (k * (unit // M))) + int(torch.randint(0, unit, (1,), generator=g, device=device))
        ph = (ph_units % unit).to(torch.float32) * (two_pi / unit)
        acc[:, 0] += torch.cos(ph)
        acc[:, 1] += torch.sin(ph)
    acc *= 0.5 / (M ** 0.5 * 3.0)  # rms ~ 0.5/3: headroom for the sum of 256 carriers
    sigma = float(0.5 / (M ** 0.5 * 3.0)) * 10 ** (-30.0 / 20.0) / 2 ** 0.5
    acc += sigma * torch.randn((N_IN, 2), device=device, dtype=torch.float32, generator=g)
    return acc.contiguous()


def bind_to_gpu_numa_node(index: int):
    """Pin this rank's host threads (and therefore its pinned-buffer pages) to the NUMA node its GPU hangs off, so the
    end-to-end leg of several ranks does not cross the socket interconnect.  Best effort: returns the node or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.lower().split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:], rest)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------- reference arm

def _load_ref():
    """The unmodified reference compiled by oracle/Makefile (perf-bench flags); None if it did not travel."""
    path = os.path.join(ROOT, "oracle", "_ref", "libdsdneo_ref_fast.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    f32p = C.POINTER(C.c_float)
    L.ref_demod_create_wideband.restype = C.c_void_p
    L.ref_demod_create_wideband.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
    L.ref_demod_destroy.argtypes = [C.c_void_p]
    L.ref_demod_block.argtypes = [C.c_void_p, f32p, C.c_int, f32p, C.c_int]
    L.simd_fir_get_impl_name.restype = C.c_char_p
    return L


def cpu_reference_run(seconds_per_thread=None, blocks_per_thread=None):
    """Times the reference's own CPU path for this workload on all host cores.

    One channel of the reference = one `struct demod_state` fed the WIDEBAND block: 8 half-band /2 stages
    (12.288 MS/s -> 48 kS/s, src/dsp/demod_pipeline.cpp:983-1001) + 135-tap channel LPF + FSK discriminator
    (full_demod, :1330-1350).  The tuner/mixer that would centre each of the 256 channels is NOT charged.
    Each thread owns one demod_state and pushes 131072-pair blocks (the reference's MAXIMUM_BUF_LENGTH).
    Whole-job IQ rate = aggregate wideband pairs/s over all threads / 256 channels.
    Falls back to the oracle port (channel LPF + discriminator only, no decimation) if oracle/_ref is absent.
    """
    import numpy as np

    cores = len(os.sched_getaffinity(0))
    L = _load_ref()
    rng = np.random.default_rng(1)
    if L is not None:
        kind = "reference"
        blk_pairs = 131072
        x = (rng.standard_normal(2 * blk_pairs) * 0.2).astype(np.float32)
        handles = [L.ref_demod_create_wideband(WIDEBAND_HZ, 8, SYMBOL_HZ, 4, 1, 0.0) for _ in range(cores)]
        outs = [np.empty(blk_pairs, np.float32) for _ in range(cores)]
        f32p = C.POINTER(C.c_float)

        def one_block(i):
            L.ref_demod_block(handles[i], x.ctypes.data_as(f32p), 2 * blk_pairs, outs[i].ctypes.data_as(f32p), blk_pairs)

        chan_div = M  # every channel consumes the whole wideband stream
        what = ("%d threads, each one reference demod_state: full_demod with 8 half-band stages (12.288 MS/s->48 kS/s) + "
                "135-tap channel LPF + FSK discriminator on 131072-pair wideband blocks, perf-bench flags, simd=%s; "
                "whole-job rate = aggregate wideband pairs/s / 256 channels (per-channel tuner mixing not charged)"
                % (cores, L.simd_fir_get_impl_name().decode()))
    else:
        kind = "port"
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import _harness as H

        blk_pairs = BLOCK_PAIRS
        O = H.oracle()
        x = (rng.standard_normal(2 * blk_pairs) * 0.2).astype(np.float32)
        chans = []
        for _ in range(cores):
            ch = H.OracleDemodChan()
            O.oracle_demod_chan_init(C.byref(ch), CHAN_HZ, 4, 1, 0.0, 1)
            chans.append(ch)
        outs = [np.empty(blk_pairs, np.float32) for _ in range(cores)]
        scr = [np.empty(2 * blk_pairs, np.float32) for _ in range(cores)]

        def one_block(i):
            O.oracle_full_demod_block(C.byref(chans[i]), H._ptr(x), 2 * blk_pairs, H._ptr(scr[i]), H._ptr(outs[i]))

        chan_div = 1.0 / M * M  # channel-rate samples: IQ rate = channel pairs/s * M / M channels
        what = ("%d threads, oracle port of channel LPF + FSK discriminator on already-channelised 8192-pair blocks "
                "(oracle/_ref absent: half-band channel selection NOT included, so this overstates the CPU)" % cores)

    counts = [0] * cores
    deadline = [0.0]

    def worker(i):
        n = 0
        if blocks_per_thread is not None:
            for _ in range(blocks_per_thread):
                one_block(i)
                n += 1
        else:
            while time.perf_counter() < deadline[0]:
                one_block(i)
                n += 1
        counts[i] = n

    for i in range(cores):
        one_block(i)  # warm-up (plans the LPF, touches buffers)
    deadline[0] = time.perf_counter() + (seconds_per_thread or 0.0)
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(i,)) for i in range(cores)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    pairs = sum(counts) * blk_pairs
    if kind == "reference":
        iq_msps = pairs / dt / M / 1e6
    else:
        iq_msps = pairs / dt / 1e6  # channel-rate pairs/s * 256 channels-per-IQ-sample / 256 channels
    if L is not None:
        for h in handles:
            L.ref_demod_destroy(h)
    return {"value": iq_msps, "unit": "MS/s", "cores": cores, "kind": kind, "sample": what, "seconds": dt,
            "blocks": sum(counts)}


def run_c2_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step_blocks = 24
    for _ in range(max(1, args.warmup)):
        cpu_reference_run(blocks_per_thread=1)
    t0 = time.perf_counter()
    vals = []
    last = None
    for _ in range(args.steps):
        last = cpu_reference_run(blocks_per_thread=per_step_blocks)
        vals.append(last["value"])
    dt = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "iq_msps", "value": value, "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "channels_at_realtime": value * 1e6 / WIDEBAND_HZ * M,
        "config": {"workload": WORKLOAD, "channels": M, "wideband_rate_hz": WIDEBAND_HZ,
                   "step": "bounded sample: %d wideband blocks of 131072 pairs per thread" % per_step_blocks},
        "cpu_baseline": {"value": value, "unit": "MS/s", "cores": last["cores"], "kind": last["kind"], "sample": last["sample"]},
        "e2e": {"value": value, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- our arm

def run_c2_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path")
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    b200 = g.load_package()
    b200.init(local)

    def barrier():
        if world > 1:
            t = torch.zeros(1, device=dev)
            dist.all_reduce(t)
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs resident in HBM (weak scaling: every rank owns one independent 256-channel band) ----
    base = make_wideband(torch, dev, seed=rank)
    bufs = [base, torch.roll(base, 777 * M + 13, 0).contiguous(), base.flip(0).contiguous()][:N_ROTATE]
    out = torch.empty((M, N_OUT), device=dev, dtype=torch.float32)
    fe = b200.Frontend(M, 8, False, WIDEBAND_HZ, BLOCK_PAIRS)
    stream = torch.cuda.current_stream(dev)

    clocks = ClockSampler(local).start()
    for i in range(args.warmup):
        fe.process_async(bufs[i % N_ROTATE], out, stream)
    fe.join(stream)
    barrier()
    b200.timing_enable(True)
    launches0 = b200.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        # pipelined C-ABI call: FIR-side kernels of step i+1 overlap the serial recurrence kernel of step i
        fe.process_async(bufs[i % N_ROTATE], out, stream)
    fe.join(stream)
    ev1.record(stream)
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = b200.launch_count() - launches0
    ktimes = b200.timing_report()
    b200.timing_enable(False)
    clk = clocks.stop()
    ms_per_step = ms_total / args.steps
    value = world * N_IN / (ms_per_step * 1e-3) / 1e6  # whole-job IQ MS/s
    # control: the same kernels one tile at a time (no overlap between the stages of consecutive tiles)
    b200.timing_enable(True)
    for i in range(min(10, args.steps)):
        fe.process_async(bufs[i % N_ROTATE], out, stream)
        fe.join(stream)
        torch.cuda.synchronize()
    kserial = {k: {"avg_ms": v["ms"] / max(1, v["launches"])} for k, v in b200.timing_report().items()}
    b200.timing_enable(False)

    # ---- roofline of the dominant kernel (algorithmic bytes per launch) ----
    alg_bytes = {
        "pfb256_kernel": 16.0 * N_IN,              # 8 B cf32 in + 8 B cf32 channel out per wideband sample
        "lpf_phase_kernel": 12.0 * M * N_OUT,      # 8 B in + 4 B phase out per channel sample
        "disc_recurrence_kernel": 8.0 * M * N_OUT  # 4 B phase in + 4 B discriminator out per channel sample
    }
    peak, peak_src = measured_hbm_peak()
    kernels = {}
    for name, rec in ktimes.items():
        avg_ms = rec["ms"] / max(1, rec["launches"])
        kernels[name] = {"launches": rec["launches"], "avg_ms": avg_ms,
                         "share": rec["ms"] / max(1e-12, sum(r["ms"] for r in ktimes.values()))}
        if name in alg_bytes:
            kernels[name]["achieved_gbs"] = alg_bytes[name] / (avg_ms * 1e-3) / 1e9
    dom = max(kernels, key=lambda k: ktimes[k]["ms"]) if kernels else None
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(dom)
    except Exception:
        pass
    roofline = None
    if dom and dom in alg_bytes:
        a = kernels[dom]["achieved_gbs"]
        roofline = {"bound": "hbm", "kernel": dom, "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
                    "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes[dom],
                    "note": "per-kernel times from CUDA events recorded by the library on the launching stream during the timed region; "
                            "the three pipeline stages of consecutive tiles run concurrently there, so a kernel's duration includes the time it "
                            "shares the SMs with the other stages' kernels (kernels_serial: the same kernels one tile at a time)",
                    "achieved_serial": (alg_bytes[dom] / (kserial[dom]["avg_ms"] * 1e-3) / 1e9) if dom in kserial else None}

    # ---- e2e: same metric through the host-buffer C-ABI (pinned host in/out, every copy inside the timed region) ----
    # Streaming form, as the reference's demod thread consumes its input ring: tile i is queued (submit_host) while the
    # consumer waits for and reads tile i-1 (wait_host), so H2D, kernels and D2H of consecutive tiles overlap.
    e2e_steps = max(3, min(args.steps, 100))

    def e2e_leg(fe_x, h_in_x, h_out_x, streaming: bool):
        def run(n):
            checksum, prev = 0.0, None
            for i in range(n):
                if streaming:
                    t = fe_x.submit_host(h_in_x[i % len(h_in_x)], h_out_x[i % 2])
                    if prev is not None:
                        fe_x.wait_host(prev)
                        checksum += float(h_out_x[(i - 1) % 2][0, 0])  # the host reads the finished tile
                    prev = t
                else:
                    fe_x.process_host(h_in_x[i % len(h_in_x)], h_out_x[0])
                    checksum += float(h_out_x[0][0, 0])
            if streaming and prev is not None:
                fe_x.wait_host(prev)
                checksum += float(h_out_x[(n - 1) % 2][0, 0])
            return checksum

        run(5)
        barrier()
        t0 = time.perf_counter()
        run(e2e_steps)
        torch.cuda.synchronize()
        ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
        barrier()
        return ms

    h_in = [torch.empty((N_IN, 2), dtype=torch.float32).pin_memory() for _ in range(N_ROTATE)]
    for i in range(N_ROTATE):
        h_in[i].copy_(bufs[i])
    h_out2 = [torch.empty((M, N_OUT), dtype=torch.float32).pin_memory() for _ in range(2)]
    h_out = h_out2[0]
    fe_h = b200.Frontend(M, 8, False, WIDEBAND_HZ, BLOCK_PAIRS)
    e2e_ms = e2e_leg(fe_h, h_in, h_out2, streaming=True)
    e2e_sync_ms = e2e_leg(fe_h, h_in, h_out2, streaming=False)
    e2e = {"value": world * N_IN / (e2e_ms * 1e-3) / 1e6, "unit": "MS/s", "h2d_bytes_per_step": N_IN * 8,
           "d2h_bytes_per_step": M * N_OUT * 4, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "timer": "host wall clock around K x {dsdneo_b200_frontend_submit_host(tile i); wait_host(tile i-1); host reads "
                    "tile i-1} + final wait, max over ranks",
           "channels_at_realtime": world * N_IN / (e2e_ms * 1e-3) / WIDEBAND_HZ * M,
           "synchronous_call": {"value": world * N_IN / (e2e_sync_ms * 1e-3) / 1e6, "ms_per_step": e2e_sync_ms,
                                "call": "dsdneo_b200_frontend_process_host (one blocking call per tile)"}}
    # the same tiles in the reference's native ingest format (cu8, widened on the GPU = simd_widen.cpp:139-147 fused
    # into the channelizer's loads): 2 B instead of 8 B per wideband sample over PCIe
    q8 = [(bufs[i] * 127.5 + 127.5).round_().clamp_(0, 255).to(torch.uint8).cpu().pin_memory() for i in range(N_ROTATE)]
    fe_q = b200.Frontend(M, 8, True, WIDEBAND_HZ, BLOCK_PAIRS)
    e2e_q_ms = e2e_leg(fe_q, q8, h_out2, streaming=True)
    e2e["cu8_input"] = {"value": world * N_IN / (e2e_q_ms * 1e-3) / 1e6, "unit": "MS/s", "ms_per_step": e2e_q_ms,
                        "h2d_bytes_per_step": N_IN * 2, "d2h_bytes_per_step": M * N_OUT * 4,
                        "note": "same workload fed as cu8 IQ (the reference's RTL ingest format); reported beside the cf32 headline"}
    fe_q.close()
    del q8

    # parity spot-check of the e2e output against the device-resident path (same state history => same bits)
    fe_a = b200.Frontend(M, 8, False, WIDEBAND_HZ, BLOCK_PAIRS)
    fe_b = b200.Frontend(M, 8, False, WIDEBAND_HZ, BLOCK_PAIRS)
    oa = fe_a.process(bufs[0]).cpu()
    ob = torch.from_numpy(fe_b.process_host(h_in[0].numpy()))
    same = bool(torch.equal(oa.view(torch.int32), ob.view(torch.int32)))

    cpu_baseline = None
    if rank == 0 and world == 1:
        cb = cpu_reference_run(seconds_per_thread=1.5)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": "iq_msps", "value": value, "unit": "MS/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "channels_at_realtime": value * 1e6 / WIDEBAND_HZ * M,
        "config": {"workload": WORKLOAD, "channels_per_gpu": M, "wideband_rate_hz": WIDEBAND_HZ, "channel_rate_hz": CHAN_HZ,
                   "samples_per_step": N_IN, "block_pairs": BLOCK_PAIRS, "blocks_per_step": N_BLOCKS,
                   "channelizer_taps_per_branch": 8, "fir_arith": "fma (reference AVX2 kernel order)",
                   "parallelism": "bands sharded over GPUs, no data-path collective",
                   "step_call": "dsdneo_b200_frontend_process_async (2-stream stage pipeline), joined before the stop event",
                   "l2_policy": "%d rotating input buffers of %.1f MB (> 126 MB L2)" % (N_ROTATE, N_IN * 8 / 1e6),
                   "host_numa_binding": "rank 0 on node %s" % numa_node if numa_node is not None else "none"},
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels,
        "kernels_serial": {"avg_ms": {k: v["avg_ms"] for k, v in kserial.items()},
                           "note": "control: one tile at a time, no overlap between the stages of consecutive tiles"},
        "cpu_baseline": cpu_baseline, "e2e_matches_device_path": same,
    }
    print(json.dumps(line))


# =============================================================================================== C3: the judged workload

C3_CH = 1024            # channels per GPU
C3_RATE = 48000
C3_PAIRS = 49152        # IQ pairs per channel per step (1.024 s)
C3_BLOCK = 8192         # one full_demod() block = the reference's DEFAULT_BUF_LENGTH (16384 floats)
C3_TILES = 5            # rotating input tiles: 5 x 49152 samples = 24576 symbols, so the rotation closes on a whole symbol
C3_SNR_DB = 20.0
C3_WORKLOAD = ("C3: 1024 synthetic P25 Phase 1 C4FM channels per GPU, cu8 IQ at 48 kS/s -> widen + full_demod -> p25_filter + "
               "getDibitSoft -> frame sync -> NID -> TSBK half-rate trellis / HDU Golay + RS(36,20,17) / LDU1,2 IMBE "
               "de-interleave + Hamming(10,6,3) + RS(24,12,13),(24,16,9) + LSD; frames, IMBE frames and dibits out")


def c3_config():
    """The SAME dictionary in both arms (driver: same_config)."""
    return {"workload": C3_WORKLOAD, "channels_per_gpu": C3_CH, "channel_rate_hz": C3_RATE, "symbol_rate_hz": 4800,
            "pairs_per_channel_per_step": C3_PAIRS, "block_pairs": C3_BLOCK, "input_format": "cu8 IQ (2 B per pair)",
            "snr_db": C3_SNR_DB,
            "traffic": "16 base channels tiled over the channels: 8 voice (HDU, 13 x {LDU1, LDU2}, TDU per 5.12 s) and 8 control "
                       "(68 three-block TSDUs per 5.12 s), valid NID / trellis / Hamming / Golay / RS / LSD coding, random payloads",
            "l2_policy": "5 rotating input tiles of 100.7 MB per GPU (503 MB > 126 MB L2); every step also streams about 0.9 GB of "
                         "intermediates through HBM",
            "parallelism": "channels sharded over GPUs (1024 per GPU), no data-path collective"}


def _c4fm_shaping_taps(ntaps=241, fs=48000.0, rs=4800.0, alpha=0.2):
    """P25 C4FM transmit shaping (TIA-102.BAAA: raised cosine alpha 0.2 times x / sin(x)), frequency-sampled."""
    import numpy as np

    N = 4096
    f = np.fft.rfftfreq(N, 1 / fs)
    T = 1 / rs
    f1, f2 = (1 - alpha) / (2 * T), (1 + alpha) / (2 * T)
    rc = np.zeros_like(f)
    rc[f <= f1] = 1.0
    m = (f > f1) & (f <= f2)
    rc[m] = 0.5 * (1 + np.cos(np.pi * T / alpha * (f[m] - f1)))
    x = np.pi * f * T
    shp = np.ones_like(f)
    shp[x > 0] = x[x > 0] / np.sin(x[x > 0])
    h = np.fft.irfft(rc * np.where(f <= f2, shp, 0.0), N)
    return np.roll(h, ntaps // 2)[:ntaps] * 10.0


def c3_base_iq(seed=0):
    """The 16 base channels as cu8 IQ, [16][5 * 49152][2]: the committed dibit streams (tests/golden/c3_p25_dibits.npz, frames
    built by the test harness encoders) C4FM-shaped, FM-modulated at +-1.8 kHz outer deviation on a CIRCULAR time axis (the
    rotation wraps without a symbol-timing jump), AWGN at C3_SNR_DB, quantised like an RTL-SDR."""
    import numpy as np

    g = np.load(os.path.join(ROOT, "tests", "golden", "c3_p25_dibits.npz"))
    dib = g["dibits"]
    levels = np.array([1.0, 3.0, -1.0, -3.0])
    taps = _c4fm_shaping_taps()
    rng = np.random.default_rng(0xC3 + seed)
    n = dib.shape[1] * 10
    assert n == C3_TILES * C3_PAIRS
    out = np.empty((dib.shape[0], n, 2), np.uint8)
    H = np.fft.rfft(np.roll(np.concatenate([taps, np.zeros(n - taps.size)]), -(taps.size // 2)))
    for c in range(dib.shape[0]):
        imp = np.zeros(n)
        imp[::10] = levels[dib[c]]
        f = np.roll(np.fft.irfft(np.fft.rfft(imp) * H, n), 9)  # 9 samples: symbol centres on the locked slicer's window
        ph = np.cumsum(f * (2 * np.pi * 600.0 / 48000.0))
        ph -= np.arange(1, n + 1) * (np.remainder(ph[-1], 2 * np.pi) / n)  # < 0.2 Hz offset: the phase closes over the rotation
        z = 0.6 * np.exp(1j * (0.3 + ph))
        sigma = 0.6 * 10 ** (-C3_SNR_DB / 20.0) / np.sqrt(2.0)
        z = z + sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        out[c, :, 0] = np.clip(np.rint(z.real * 127.5 + 127.5), 0, 255)
        out[c, :, 1] = np.clip(np.rint(z.imag * 127.5 + 127.5), 0, 255)
    return out


def _p25_filter_taps():
    """The reference's normalised p25_filter taps at 10 samples per symbol (read out of the unmodified reference by
    tests/golden/make_golden.py; the dsd-neo glue passes its own table)."""
    import numpy as np

    g = np.load(os.path.join(ROOT, "tests", "golden", "sps_fir_taps.npz"))
    return np.ascontiguousarray(g["f0_sps10"], np.float32)  # filter 0 = p25_filter


# ----------------------------------------------------------------------------------------------- C3 reference arm

class RefPool:
    """The UNMODIFIED reference on the host cores, one forked worker per core (the reference keeps process-global state:
    one channel per process).  Every worker owns one channel of the same synthetic workload and runs, per tile of 49152 cu8 pairs:
    widen_u8_to_f32_bias127 -> full_demod (6 blocks of 8192 pairs, perf-bench flags, AVX2 dispatch) -> then per step
    getDibitSoft over the step's discriminator samples through the reference's own hook seam -> frame sync search ->
    dsd_dispatch_handle_p25p1 (NID, TSBK / HDU / LDU handlers, FEC leaves) on every sync.  States are created once; a step
    is `tiles` tiles per worker, so steps measure steady state."""

    def __init__(self, base_u8, cores=None):
        import multiprocessing as mp

        self.cores = cores or len(os.sched_getaffinity(0))
        self.ctx = mp.get_context("fork")
        self.base = base_u8
        self.tiles = self.ctx.Value("i", 1)
        self.stop = self.ctx.Value("i", 0)
        self.b0 = self.ctx.Barrier(self.cores + 1)
        self.b1 = self.ctx.Barrier(self.cores + 1)
        self.frames = self.ctx.Array("l", self.cores)
        self.kind = None
        fast = os.path.join(ROOT, "oracle", "_ref", "libdsdneo_ref_fast.so")
        p25 = os.path.join(ROOT, "oracle", "_ref", "libdsdneo_ref_p25.so")
        if not (os.path.exists(fast) and os.path.exists(p25)):
            raise SystemExit("bench.py: oracle/_ref is missing (run python __graft_entry__.py where /root/reference exists)")
        self.paths = (fast, p25)
        self.procs = [self.ctx.Process(target=self._worker, args=(i,), daemon=True) for i in range(self.cores)]
        for p in self.procs:
            p.start()

    def _worker(self, idx):
        import numpy as np

        try:
            os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[idx % len(os.sched_getaffinity(0))]})
        except Exception:
            pass
        L, P = C.CDLL(self.paths[0]), C.CDLL(self.paths[1])
        f32p, u8p = C.POINTER(C.c_float), C.POINTER(C.c_ubyte)
        L.ref_demod_create.restype = C.c_void_p
        L.ref_demod_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
        L.ref_demod_block.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.widen_u8_to_f32_bias127.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.ref_sym_create.restype = C.c_void_p
        L.ref_sym_create.argtypes = [C.c_int] * 7
        L.ref_sym_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
        L.ref_sym_get_dibits.restype = C.c_long
        L.ref_sym_get_dibits.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        P.ref_p25_decode_frame.restype = C.c_long
        P.ref_p25_decode_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_void_p]
        rec = (C.c_ubyte * P.ref_p25_frame_size())()
        demod = L.ref_demod_create(C3_RATE, 4800, 4, 1, 0.0)  # DSD_CH_LPF_PROFILE_P25_C4FM
        sym = L.ref_sym_create(C3_RATE, 4800, 0, 0, 1, 128, 1024)
        u8 = self.base[idx % self.base.shape[0]]
        f32 = np.empty(2 * C3_PAIRS, np.float32)
        sync = np.array([int(c) for c in "111113113311333313133333"], np.uint8)
        tile_no = 0
        disc = dib = rel = llr = symv = None
        while True:
            self.b0.wait()
            if self.stop.value:
                break
            k_tiles = self.tiles.value
            if disc is None or disc.size != k_tiles * C3_PAIRS:
                disc = np.empty(k_tiles * C3_PAIRS, np.float32)
                cap = disc.size // 9 + 16
                dib, rel = np.empty(cap, np.uint8), np.empty(cap, np.uint8)
                llr, symv = np.empty(2 * cap, np.int16), np.empty(cap, np.float32)
            for k in range(k_tiles):
                t = tile_no % C3_TILES
                tile_no += 1
                src = u8[t * C3_PAIRS:(t + 1) * C3_PAIRS]
                L.widen_u8_to_f32_bias127(src.ctypes.data, f32.ctypes.data, 2 * C3_PAIRS)
                for b in range(C3_PAIRS // C3_BLOCK):
                    L.ref_demod_block(demod, f32.ctypes.data + 8 * b * C3_BLOCK, 2 * C3_BLOCK,
                                      disc.ctypes.data + 4 * (k * C3_PAIRS + b * C3_BLOCK), C3_BLOCK)
            L.ref_sym_feed(sym, disc.ctypes.data, disc.size)
            nd = L.ref_sym_get_dibits(sym, dib.size, 12, dib.ctypes.data, rel.ctypes.data, llr.ctypes.data, symv.ctypes.data)
            d = dib[:nd]
            hits = np.nonzero((np.lib.stride_tricks.sliding_window_view(d, 24) == sync).all(axis=1))[0] + 23
            n_fr = 0
            for p in hits:
                if p + 864 < nd:
                    P.ref_p25_decode_frame(dib.ctypes.data, rel.ctypes.data, llr.ctypes.data, nd, int(p) + 1, 0, rec)
                    n_fr += 1
            self.frames[idx] = n_fr
            self.b1.wait()

    def step(self, tiles):
        self.tiles.value = tiles
        t0 = time.perf_counter()
        self.b0.wait()
        self.b1.wait()
        return time.perf_counter() - t0

    def close(self):
        self.stop.value = 1
        try:
            self.b0.wait(timeout=5)
        except Exception:
            pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()

    def describe(self, tiles):
        return ("%d forked workers (one per core), each ONE channel of the same synthetic workload through the unmodified reference: "
                "widen_u8_to_f32_bias127 + full_demod (perf-bench flags, AVX2 FIR) + getDibitSoft (p25_filter, hook seam) + frame "
                "sync + dsd_dispatch_handle_p25p1 (NID / TSBK / HDU / LDU handlers); states created once, %d tiles of 1.024 s per "
                "worker per step" % (self.cores, tiles))


def c3_reference_measure(steps, warmup, target_step_s=None):
    if target_step_s is None:  # about 1.2 s of CPU work per step, shortened so that the whole arm ends within a few minutes
        target_step_s = min(1.2, max(0.3, 150.0 / max(1, steps + max(0, warmup))))
    base = c3_base_iq()
    pool = RefPool(base)
    try:
        pool.step(1)  # plans the LPF, touches buffers
        t1 = pool.step(2) / 2
        tiles = int(max(4, min(256, round(target_step_s / max(t1, 1e-4)))))
        for _ in range(max(0, warmup)):
            pool.step(tiles)
        times = [pool.step(tiles) for _ in range(steps)]
        frames = sum(pool.frames[:])
    finally:
        pool.close()
    dt = sum(times) / len(times)
    value = pool.cores * tiles * C3_PAIRS / dt / 1e6
    return {"value": value, "unit": "MS/s", "cores": pool.cores, "kind": "reference", "sample": pool.describe(tiles),
            "ms_per_step": dt * 1e3, "tiles_per_worker_per_step": tiles, "frames_per_step": int(frames)}


def run_c3_reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    m = c3_reference_measure(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "iq_msps", "value": m["value"], "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "channels_at_realtime": m["value"] * 1e6 / C3_RATE, "config": c3_config(),
        "cpu_baseline": {k: m[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": m["value"], "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "step": "bounded sample: %d tiles of 49152 pairs per worker per step, %d frames decoded in the last step"
                                   % (m["tiles_per_worker_per_step"], m["frames_per_step"]),
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- C3 our arm

def run_c3_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path")
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None  # before any pinned allocation (first touch)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    b200 = g.load_package()
    b200.init(local)

    def barrier():
        if world > 1:
            t = torch.zeros(1, device=dev)
            dist.all_reduce(t)
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs: this rank's 1024 channels (channel c of the job = base channel c % 16, rotated per rank) ----
    base = c3_base_iq(seed=rank)
    idx = (np.arange(C3_CH) + 5 * rank) % base.shape[0]
    h_tiles = []
    for t in range(C3_TILES):
        ht = torch.empty((C3_CH, C3_PAIRS, 2), dtype=torch.uint8).pin_memory()
        ht.copy_(torch.from_numpy(np.ascontiguousarray(base[idx, t * C3_PAIRS:(t + 1) * C3_PAIRS])))
        h_tiles.append(ht)
    d_tiles = [t.to(dev) for t in h_tiles]
    taps = _p25_filter_taps()
    watch = int(getattr(args, "auto_reacquire", 0))  # developer switch: the loss-of-sync watch on the device (default off)
    rx = b200.P25p1Rx(C3_CH, taps, rate_hz=C3_RATE, block_pairs=C3_BLOCK, max_pairs_per_call=C3_PAIRS, input_cu8=True, max_hits=32,
                      auto_reacquire_tiles=watch)
    out = rx.alloc_device_out(dev)
    stream = torch.cuda.current_stream(dev)

    clocks = ClockSampler(local).start()
    n_warm = max(args.warmup, C3_TILES)  # one full rotation: the threshold trackers settle, every code path is loaded
    # Tiles are submitted to the bank's three-stage pipeline (channel filter | recurrences + matched filter | slicer + frames):
    # consecutive tiles overlap on the device, every tile runs every stage, results complete in submission order.
    last = -1
    for i in range(n_warm):
        last = rx.submit(d_tiles[i % C3_TILES], C3_PAIRS, out, stream)
    rx.wait(last, stream)
    barrier()
    b200.timing_enable(True)
    launches0 = b200.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_frames = n_voice = n_good = 0
    ev0.record(stream)
    for i in range(args.steps):
        last = rx.submit(d_tiles[(i + n_warm) % C3_TILES], C3_PAIRS, out, stream)
    rx.wait(last, stream)
    ev1.record(stream)
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = b200.launch_count() - launches0
    ktimes = b200.timing_report()
    # the same kernels one tile at a time (process = submit + wait): what each costs when it has the GPU to itself
    n_serial = min(args.steps, 10)
    b200.timing_enable(True)
    evs0, evs1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs0.record(stream)
    for i in range(n_serial):
        rx.process(d_tiles[(i + n_warm + args.steps) % C3_TILES], C3_PAIRS, out, stream)
    evs1.record(stream)
    torch.cuda.synchronize()
    serial_ms_per_step = evs0.elapsed_time(evs1) / n_serial
    kserial = {k: {"avg_ms": v["ms"] / max(1, v["launches"])} for k, v in b200.timing_report().items()}
    b200.timing_enable(False)
    clk = clocks.stop()
    fr, vo = rx.records(out)  # the last step's records: how much real traffic the step decoded
    n_frames, n_voice = int(fr.size), int(vo.size)
    n_good = int(((fr["nid_status"] > 0) & (((fr["duid"] == 7) & (fr["n_tsbk"] > 0)) | ((fr["rs_kind"] > 0) & (fr["rs_status"] < 2))
                                           | (fr["duid"] == 3))).sum())
    n_sym = int(out["counts"].sum().item())
    ms_per_step = ms_total / args.steps
    value = world * C3_CH * C3_PAIRS / (ms_per_step * 1e-3) / 1e6

    # ---- roofline of the dominant kernel: algorithmic bytes per launch ----
    n_s = float(C3_CH * C3_PAIRS)
    alg_bytes = {
        "lpf_phase_kernel": 6.0 * n_s,               # 2 B cu8 in (widened in the staging) + 4 B phase out per pair
        "disc_recurrence_kernel": 8.0 * n_s,         # 4 B in + 4 B out
        "sps_fir_kernel": 8.0 * n_s,                 # 4 B in + 4 B out per sample
        "symbolize_kernel": 4.0 * n_s + 10.0 * n_sym,  # sps x 4 B in + 10 B out per symbol (the reference's .bin record)
        "frame_sync_search_kernel": 4.0 * n_sym,
    }
    peak, peak_src = measured_hbm_peak()
    tot = sum(r["ms"] for r in ktimes.values()) or 1e-12
    kernels = {}
    for name, rec in ktimes.items():
        avg_ms = rec["ms"] / max(1, rec["launches"])
        per_step = rec["ms"] / args.steps
        kernels[name] = {"launches": rec["launches"], "avg_ms": avg_ms, "ms_per_step": per_step, "share": rec["ms"] / tot}
        if name in alg_bytes:
            kernels[name]["achieved_gbs"] = alg_bytes[name] / (avg_ms * 1e-3) / 1e9
    dom = max((k for k in kernels if k in alg_bytes), key=lambda k: ktimes[k]["ms"]) if kernels else None
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(dom)
    except Exception:
        pass
    roofline = None
    if dom:
        a = kernels[dom]["achieved_gbs"]
        roofline = {"bound": "hbm", "kernel": dom, "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes[dom],
                    "note": "per-kernel times from CUDA events recorded by the library on the launching stream during the timed region"}

    # ---- e2e: pinned host cu8 IQ in, host frames / IMBE frames / dibits out, every copy inside the timed region ----
    e2e_steps = max(3, min(args.steps, 100))
    rx_h = b200.P25p1Rx(C3_CH, taps, rate_hz=C3_RATE, block_pairs=C3_BLOCK, max_pairs_per_call=C3_PAIRS, input_cu8=True, max_hits=32,
                        auto_reacquire_tiles=watch)
    DEPTH = 6  # host tiles in flight = DEPTH - 1: H2D, the bank's four pipeline stages, D2H (at most six tickets outstanding)
    h_outs = [rx_h.alloc_host_out() for _ in range(DEPTH)]
    d2h = [0]
    seq = [0]  # tiles are fed in rotation order across calls of run(): the stream stays continuous (the locked slicer does not
               # re-acquire symbol timing, so a repeated or skipped tile would slip it by a fraction of a symbol)

    def run(n, count=False):
        chk = 0
        pending = []  # (ticket, tile number)

        def finish(entry):
            nonlocal chk
            t, k = entry
            rx_h.wait_host(t)
            o = h_outs[k % DEPTH]
            chk += int(o["totals"][0]) + int(o["dibits"][0, 0])  # the host reads the finished tile
            if count:
                d2h[0] += int(o["totals"][0]) * 128 + int(o["totals"][1]) * 1944

        for _ in range(n):
            k = seq[0]
            seq[0] += 1
            if len(pending) == DEPTH - 1:
                finish(pending.pop(0))
            pending.append((rx_h.submit_host(h_tiles[k % C3_TILES], C3_PAIRS, h_outs[k % DEPTH]), k))
        while pending:
            finish(pending.pop(0))
        return chk

    run(C3_TILES + 1)
    barrier()
    t0 = time.perf_counter()
    run(e2e_steps, count=True)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    barrier()
    fixed_d2h = C3_CH * rx_h.dibit_pitch + C3_CH * 4 + 8
    e2e = {"value": world * C3_CH * C3_PAIRS / (e2e_ms * 1e-3) / 1e6, "unit": "MS/s", "h2d_bytes_per_step": C3_CH * C3_PAIRS * 2,
           "d2h_bytes_per_step": int(fixed_d2h + d2h[0] / e2e_steps), "ms_per_step": e2e_ms, "steps": e2e_steps,
           "channels_at_realtime": world * C3_CH * C3_PAIRS / (e2e_ms * 1e-3) / C3_RATE,
           "timer": "host wall clock around K x {wait_host(tile i-5); host reads tile i-5; dsdneo_b200_p25p1_rx_submit_host(tile i)} "
                    "+ final waits, max over ranks (five tiles in flight)",
           "d2h_contents": "dibit stream + counts + frame records + IMBE frame records"}
    # device path == host path (same state history => same bytes)
    fr_h, vo_h = rx_h.host_records(h_outs[(seq[0] - 1) % DEPTH])

    cpu_baseline = None
    if rank == 0 and world == 1:
        try:
            m = c3_reference_measure(steps=3, warmup=1, target_step_s=2.0)
            cpu_baseline = {k: m[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except SystemExit as e:
            cpu_baseline = {"value": None, "unit": "MS/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": "iq_msps", "value": value, "unit": "MS/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, C3_TILES),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "channels_at_realtime": value * 1e6 / C3_RATE, "config": c3_config(),
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels,
        "kernels_serial": {"ms_per_step": serial_ms_per_step, "steps": n_serial,
                           "avg_ms": {k: v["avg_ms"] for k, v in kserial.items()},
                           "note": "control: one tile at a time (dsdneo_b200_p25p1_rx_process), no overlap between the pipeline stages"},
        "cpu_baseline": cpu_baseline,
        "step_detail": {"call": "dsdneo_b200_p25p1_rx_submit + _wait (device buffers) / _submit_host + _wait_host (host buffers)",
                        "symbols_per_step": n_sym, "frames_per_step": n_frames, "frames_decoded_ok": n_good,
                        "imbe_frames_per_step": 9 * n_voice, "host_numa_binding": ("node %s" % numa_node) if numa_node is not None else
                        ("none (one rank)" if world == 1 else "none (the host exposes no NUMA node for the GPU)"),
                        "e2e_frames_last_step": int(fr_h.size), "e2e_voice_last_step": int(vo_h.size),
                        "auto_reacquire_tiles": watch},
    }
    print(json.dumps(line))


def run_c3_one_stream(args):
    """North-star multi-GPU shape at the size it is meant for (BASELINE configs C4 / C5): ONE wideband
    cu8 stream of 1024 x N channels (48 kS/s each: 49.152 x N MS/s), the raw tile reaches every GPU through the path's single
    collective, every GPU runs the bin-pruned polyphase channelizer for its channel class k = rank (mod N) and the C3 receive
    bank on those 1024 channels.  Per-GPU work is fixed as N grows (weak scaling); the wideband stream grows with N.
    `value`: the tile resident on the ingest rank's GPU, NCCL broadcast inside the timed region.  `e2e`: every rank ingests
    1/N of each tile from its own pinned host memory (H2D), one NCCL all-gather assembles the tile on every GPU, records and
    dibits return to the host."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path")
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    b200 = g.load_package()
    b200.init(local)
    from dsdneo_b200 import shard

    def barrier():
        if world > 1:
            t = torch.zeros(1, device=dev)
            dist.all_reduce(t)
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    Mw = C3_CH * world
    T = 8
    taps = _p25_filter_taps()
    sr = shard.ShardedP25Rx(b200, Mw, rank, world, taps, C3_PAIRS, rate_hz=C3_RATE, block_pairs=C3_BLOCK, taps_per_branch=T, device=dev,
                            channels_cu8=args.channels_cu8)
    n_tile = C3_PAIRS * Mw  # wideband samples per tile
    n_slice = n_tile // world
    # ---- the wideband stream: synthesised on every rank's GPU from the same 16 base channels (deterministic), so the ingest
    # rank holds the whole tiles and every rank can keep its own 1/N slices in pinned host memory for the e2e leg ----
    base = c3_base_iq(seed=0)
    bc = torch.from_numpy(((base.astype(np.float32) - 127.5) / 127.5)).to(dev)
    # synthesis + analysis bank delay the channels by T - 1 channel samples; the source is advanced by as much (circular axis) so
    # the symbol centres stay where the bank's locked slicer samples them, as in the per-channel C3 input
    bc = torch.roll(torch.complex(bc[..., 0], bc[..., 1]), -(T - 1), dims=1).contiguous()
    chan_of_bin = (torch.arange(Mw, device=dev) * 7 + 3) % base.shape[0]
    wide = shard.synthesize_wideband(torch, bc, chan_of_bin, Mw, sr.cz.prototype(), T)
    del bc
    # the rotation is circular: the samples before tile 0 are the end of tile 4, so the stream starts without the channelizer's
    # start-up ramp (the bank's slicer is not re-acquired after start-up: DESIGN.md section 8 item 1)
    sr.cz.prime(wide[-(T - 1) * Mw:].contiguous())
    h_slices = []
    for t in range(C3_TILES):
        hs = torch.empty((n_slice, 2), dtype=torch.uint8).pin_memory()
        hs.copy_(wide[t * n_tile + rank * n_slice:t * n_tile + (rank + 1) * n_slice])
        h_slices.append(hs)
    if rank == 0:
        d_tiles = [wide[t * n_tile:(t + 1) * n_tile] for t in range(C3_TILES)]
    else:
        del wide
        d_tiles = [None] * C3_TILES
    torch.cuda.empty_cache()
    out = sr.rx.alloc_device_out(dev)
    stream = torch.cuda.current_stream(dev)

    clocks = ClockSampler(local).start()
    n_warm = max(args.warmup, C3_TILES)
    seq = 0
    sr.distribute(d_tiles[0], "broadcast")
    last = -1
    for i in range(n_warm):
        seq += 1
        sr.distribute(d_tiles[seq % C3_TILES], "broadcast")  # tile i + 1 travels under the kernels of tile i
        last = sr.submit(out, stream)
    sr.rx.wait(last, stream)
    barrier()
    b200.timing_enable(True)
    launches0 = b200.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        seq += 1
        sr.distribute(d_tiles[seq % C3_TILES], "broadcast")
        last = sr.submit(out, stream)
    sr.rx.wait(last, stream)
    ev1.record(stream)
    barrier()
    ms_per_step = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    launches = b200.launch_count() - launches0
    ktimes = b200.timing_report()
    b200.timing_enable(False)
    fr, vo = sr.rx.records(out)
    n_frames, n_voice = int(fr.size), int(vo.size)
    n_good = int(((fr["nid_status"] > 0) & (((fr["duid"] == 7) & (fr["n_tsbk"] > 0)) | ((fr["rs_kind"] > 0) & (fr["rs_status"] < 2))
                                           | (fr["duid"] == 3))).sum())
    value = Mw * C3_PAIRS / (ms_per_step * 1e-3) / 1e6
    # drain the tile that was distributed but not submitted, so the e2e leg starts in step
    last = sr.submit(out, stream)
    sr.rx.wait(last, stream)
    torch.cuda.synchronize()

    # ---- e2e: every rank feeds 1/N of each tile from pinned host memory; all-gather; records + dibits back to the host ----
    e2e_steps = max(3, min(args.steps, 60))
    DEPTH = 5
    outs = [sr.rx.alloc_device_out(dev) for _ in range(DEPTH)]
    h_outs = [sr.rx.alloc_host_out() for _ in range(DEPTH)]
    d2h_stream = torch.cuda.Stream(device=dev)
    d2h_records = torch.cuda.Stream(device=dev)
    done = [None] * DEPTH
    d2h_bytes = [0]

    def finish(k):
        o, h = outs[k % DEPTH], h_outs[k % DEPTH]
        done[k % DEPTH].synchronize()
        nf, nv = int(h["totals"][0]), int(h["totals"][1])
        with torch.cuda.stream(d2h_records):  # its own stream: the other copy stream is already waiting for a later tile
            h["frames"][:nf].copy_(o["frames"][:nf], non_blocking=True)
            h["voices"][:nv].copy_(o["voices"][:nv], non_blocking=True)
        d2h_records.synchronize()
        d2h_bytes[0] += nf * 128 + nv * 1944
        return nf + int(h["dibits"][0, 0])

    def run(n):
        nonlocal seq
        chk = 0
        sr.distribute(h_slices[(seq + 1) % C3_TILES], "allgather")
        for k in range(n):
            seq += 1
            if k + 1 < n:
                sr.distribute(h_slices[(seq + 1) % C3_TILES], "allgather")
            o, h = outs[k % DEPTH], h_outs[k % DEPTH]
            t = sr.submit(o, stream)
            sr.rx.wait(t, d2h_stream)  # only the copy stream waits for the tile: the next tile's kernels follow at once
            with torch.cuda.stream(d2h_stream):
                h["totals"].copy_(o["totals"], non_blocking=True)
                h["counts"].copy_(o["counts"], non_blocking=True)
                h["dibits"].copy_(o["dibits"], non_blocking=True)
                e = torch.cuda.Event()
                e.record(d2h_stream)
            done[k % DEPTH] = e
            if k >= DEPTH - 1:
                chk += finish(k - (DEPTH - 1))
        for k in range(max(0, n - (DEPTH - 1)), n):
            chk += finish(k)
        return chk

    run(C3_TILES)
    barrier()
    d2h_bytes[0] = 0
    t0 = time.perf_counter()
    run(e2e_steps)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    barrier()
    clk = clocks.stop()
    fixed_d2h = sr.n_local * sr.rx.dibit_pitch + sr.n_local * 4 + 8
    e2e = {"value": Mw * C3_PAIRS / (e2e_ms * 1e-3) / 1e6, "unit": "MS/s", "h2d_bytes_per_step": n_slice * 2 * world,
           "d2h_bytes_per_step": int((fixed_d2h + d2h_bytes[0] / e2e_steps) * world), "ms_per_step": e2e_ms, "steps": e2e_steps,
           "channels_at_realtime": Mw * C3_PAIRS / (e2e_ms * 1e-3) / C3_RATE,
           "timer": "host wall clock, max over ranks; per step and rank: H2D of the rank's 1/N slice of the tile (pinned host), "
                    "NCCL all-gather, channelizer, receive bank, D2H of dibits + counts + frame / IMBE frame records"}
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    tot = sum(r["ms"] for r in ktimes.values()) or 1e-12
    kernels = {name: {"launches": rec["launches"], "avg_ms": rec["ms"] / max(1, rec["launches"]), "share": rec["ms"] / tot}
               for name, rec in ktimes.items()}
    peak, peak_src = measured_hbm_peak()
    roofline = None
    if "pfbn_kernel" in kernels or "pfb256_kernel" in kernels:
        kn = "pfbn_kernel" if "pfbn_kernel" in kernels else "pfb256_kernel"
        algb = n_tile * 2.0 + sr.n_local * C3_PAIRS * (2.0 if args.channels_cu8 else 8.0)  # the whole cu8 tile in, this rank's channels out
        a = algb / (kernels[kn]["avg_ms"] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": kn, "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": algb}
    cfg = c3_config()
    cfg["workload"] = ("C3 receive banks behind ONE wideband stream: %d channels x 48 kS/s = %.3f MS/s of cu8 IQ, bin-pruned polyphase "
                       "channelizer (8 taps per branch) + %s" % (Mw, Mw * C3_RATE / 1e6, C3_WORKLOAD))
    cfg["input_format"] = "one wideband cu8 IQ stream (2 B per sample), %d bytes per 1.024 s tile" % (n_tile * 2)
    cfg["parallelism"] = ("channel class k = rank (mod N) per GPU; the only collective is one NCCL broadcast (device-resident leg) / "
                          "all-gather (e2e leg, every rank ingests 1/N of the tile) of the raw tile per step, on a side stream")
    cfg["l2_policy"] = "5 rotating wideband tiles of %.1f MB" % (n_tile * 2 / 1e6)
    cfg["channel_format"] = "cu8 (re-quantised by the channelizer, gain sqrt(M))" if args.channels_cu8 else "cf32"
    print(json.dumps({
        "metric": "iq_msps", "value": value, "unit": "MS/s", "n_gpus": world, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "channels_at_realtime": value * 1e6 / C3_RATE, "config": cfg, "clocks": clk, "e2e": e2e,
        "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels,
        "step_detail": {"frames_per_step_rank0": n_frames, "frames_decoded_ok_rank0": n_good, "imbe_frames_per_step_rank0": 9 * n_voice,
                        "host_numa_binding": ("node %s" % numa_node) if numa_node is not None else "none"}}))


def run_channel_sharded(args):
    """North-star multi-GPU shape: ONE wideband stream, one NCCL broadcast of each raw IQ tile
    from rank 0, every rank runs the channelizer and demodulates its own contiguous channel range.  Strong scaling of a
    fixed band; not the driver's default line (that is --shard bands)."""
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    b200 = g.load_package()
    b200.init(local)
    from dsdneo_b200 import shard

    if args.channels != M:
        raise SystemExit("bench.py --shard channels: the channelizer kernel is built for %d channels" % M)
    base = make_wideband(torch, dev, seed=0)
    bufs = [base, torch.roll(base, 777 * M + 13, 0).contiguous(), base.flip(0).contiguous()]
    if rank != 0:
        for b in bufs:
            b.zero_()  # only the ingest rank holds data; everyone else receives it through the broadcast
    sf = shard.ShardedFrontend(b200, M, rank, world, 8, False, WIDEBAND_HZ, BLOCK_PAIRS)
    out = torch.empty((sf.hi - sf.lo, N_OUT), device=dev, dtype=torch.float32)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            t = torch.zeros(1, device=dev)
            dist.all_reduce(t)
        torch.cuda.synchronize()

    clocks = ClockSampler(local).start()
    for i in range(max(3, args.warmup)):
        sf.process(bufs[i % 3], out)
    barrier()
    launches0 = b200.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    sf.prefetch(bufs[0])
    for i in range(args.steps):
        cur = bufs[i % 3]
        sf.process(cur, out)
        if i + 1 < args.steps:
            sf.prefetch(bufs[(i + 1) % 3])  # broadcast of tile i+1 on a side stream, under the kernels of tile i
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = b200.launch_count() - launches0
    clk = clocks.stop()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    ms_per_step = ms / args.steps
    value = N_IN / (ms_per_step * 1e-3) / 1e6
    print(json.dumps({
        "metric": "iq_msps", "value": value, "unit": "MS/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "channels_at_realtime": value * 1e6 / WIDEBAND_HZ * M,
        "config": {"workload": WORKLOAD + " -- ONE band, channel ranges sharded over GPUs", "channels": M,
                   "channels_per_gpu": [shard.channel_range(r, world, M)[1] - shard.channel_range(r, world, M)[0] for r in range(world)],
                   "samples_per_step": N_IN, "collective": "one NCCL broadcast of the raw cf32 tile per step (%d bytes), "
                   "double-buffered on a side stream" % (N_IN * 8),
                   "l2_policy": "3 rotating input buffers of %.1f MB" % (N_IN * 8 / 1e6)},
        "clocks": clk, "gpu_launches": int(launches)}))


# ----------------------------------------------------------------------------------------------- optional workload: CQPSK

def cqpsk_cpu_reference(seconds=6.0, rate=24000, sps=5):
    """The reference's own CQPSK full_demod() (channel LPF + AGC / FLL / Gardner / Costas, output_kind SYMBOL_CQPSK) on all
    host cores: one `struct demod_state` per thread fed 100 ms blocks of one synthetic channel (perf-bench flags)."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _harness as H

    path = os.path.join(ROOT, "oracle", "_ref", "libdsdneo_ref_fast.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    f32p = C.POINTER(C.c_float)
    L.ref_demod_create_cqpsk.restype = C.c_void_p
    L.ref_demod_create_cqpsk.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
    L.ref_demod_block.argtypes = [C.c_void_p, f32p, C.c_int, f32p, C.c_int]
    L.ref_demod_destroy.argtypes = [C.c_void_p]
    cores = len(os.sched_getaffinity(0))
    bp = rate // 10
    rng = np.random.default_rng(3)
    x = np.ascontiguousarray(H.synth_cqpsk_iq(rng, 10 * bp // sps + 2, sps=sps, snr_db=18.0, cfo=0.004)[0][:10 * bp]).reshape(-1)
    handles = [L.ref_demod_create_cqpsk(rate, rate // sps, sps, 1, 0.0, 0.0, 0) for _ in range(cores)]
    outs = [np.empty(bp + 2, np.float32) for _ in range(cores)]
    counts = [0] * cores
    deadline = [0.0]

    def worker(i):
        n_sym, b = 0, 0
        while time.perf_counter() < deadline[0]:
            blk = x[(b % 10) * 2 * bp:(b % 10 + 1) * 2 * bp]
            n_sym += L.ref_demod_block(handles[i], blk.ctypes.data_as(f32p), 2 * bp, outs[i].ctypes.data_as(f32p), bp + 2)
            b += 1
        counts[i] = n_sym

    deadline[0] = time.perf_counter() + seconds
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(i,)) for i in range(cores)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    for h in handles:
        L.ref_demod_destroy(h)
    return {"value": sum(counts) / dt / 1e6, "unit": "Msym/s", "cores": cores, "kind": "reference",
            "sample": "%d threads x %.0f s, each one reference demod_state in CQPSK symbol mode (full_demod: channel LPF + AGC + "
                      "FLL + Gardner + diff phasor + Costas) on 100 ms blocks at %d S/s, sps %d, perf-bench flags" % (cores, seconds, rate, sps)}


def run_cqpsk_workload(args):
    """Developer line (NOT the judged C2 line): N CQPSK channels, IQ -> symbols -> dibits on one GPU, with the reference's CPU
    path beside it.  Same JSON keys as the main line where they apply."""
    import numpy as np
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _harness as H
    import __graft_entry__ as g

    b200 = g.load_package()
    b200.init(0)
    n_ch, rate, sps = args.channels if args.channels != M else 1024, 24000, 5
    bp, nb = rate // 10, 10
    rng = np.random.default_rng(5)
    base = [H.synth_cqpsk_iq(rng, bp * nb // sps + 2, sps=sps, snr_db=18.0, cfo=0.004 * (c - 4), timing=0.11 * c)[0][:bp * nb]
            for c in range(8)]
    h_x = torch.from_numpy(np.stack([base[c % 8] for c in range(n_ch)])).pin_memory()
    x = h_x.cuda()
    bank = b200.CqpskBank(n_ch, rate, ted_sps=[sps] * n_ch)
    slicer = b200.CqpskSlicer(n_ch)

    def step(src):
        sym, counts = bank.full_demod(src, bp, nb)
        return slicer.run(sym, counts.sum(dim=1, dtype=torch.int32).contiguous()), counts

    for _ in range(max(3, args.warmup)):
        res, counts = step(x)
    torch.cuda.synchronize()
    n_sym = int(counts.sum().item())
    b200.timing_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = b200.launch_count()
    e0.record()
    for _ in range(args.steps):
        res, counts = step(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = b200.launch_count() - launches0
    rep = b200.timing_report()
    b200.timing_enable(False)
    # end to end: pinned host IQ in, host dibits out, every step
    h_out = torch.empty((n_ch, res["dibits"].shape[1]), dtype=torch.uint8).pin_memory()
    t0 = time.perf_counter()
    k_e2e = max(3, args.steps // 4)
    for _ in range(k_e2e):
        d = h_x.cuda(non_blocking=True)
        r, _ = step(d)
        h_out.copy_(r["dibits"], non_blocking=True)
        torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / k_e2e * 1e3
    kern = {k: v["ms"] / v["launches"] for k, v in rep.items()}
    chain_ms = kern.get("cqpsk_chain_kernel", 0.0)
    alg = n_ch * bp * nb * 8.0 + n_sym * 4.0
    peak, peak_src = measured_hbm_peak()
    cpu = cqpsk_cpu_reference()
    print(json.dumps({
        "metric": "cqpsk_channel_msym_per_s", "value": n_sym / ms / 1e3, "unit": "Msym/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "channels_at_realtime": n_sym / (ms * 1e-3) / (rate / sps),
        "config": {"workload": "CQPSK (section 8f rank 3, developer line): %d synthetic P25 LSM-like channels at %d S/s, sps %d, "
                               "1 s of signal per step: channel LPF -> AGC/FLL/Gardner/Costas -> symbol-rate slicer" % (n_ch, rate, sps),
                   "channels": n_ch, "l2_policy": "input %.0f MB per step (> 126 MB L2)" % (n_ch * bp * nb * 8 / 1e6)},
        "e2e": {"value": n_sym / e2e_ms / 1e3, "unit": "Msym/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": n_ch * bp * nb * 8,
                "d2h_bytes_per_step": int(h_out.numel())},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "cqpsk_chain_kernel", "achieved": alg / (chain_ms * 1e-3) / 1e9 if chain_ms else None,
                     "peak": peak, "unit": "GB/s", "frac": (alg / (chain_ms * 1e-3) / 1e9 / peak) if chain_ms else None,
                     "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                     "note": "latency / issue bound by the reference's feedback loops (DESIGN 4.3e); HBM fraction low by construction"},
        "kernels": kern, "cpu_baseline": cpu}))


# ----------------------------------------------------------------------------------------------- optional workload: FEC leaves

def run_fec_workload(args):
    """Developer line (NOT the judged C2 line): the batched FEC leaves on device-resident inputs (CUDA events) with the
    reference's own functions timed natively beside them on ONE host core (they keep process-global state).
    Inputs: valid codewords (the all-zero word of each linear code; encoded trellis paths) with a few channel errors."""
    import numpy as np
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _harness as H
    import __graft_entry__ as g

    b200 = g.load_package()
    b200.init(0)
    lib = b200.lib()
    rng = np.random.default_rng(11)
    N = 1 << 17
    n_cpu = 4096

    def flips(shape_n, width, k):
        e = np.zeros((shape_n, width), np.uint8)
        for _ in range(k):
            e[np.arange(shape_n), rng.integers(0, width, shape_n)] ^= 1
        return e

    items = {}
    # Golay(24,12): zero codeword + 2 errors
    gol = flips(N, 24, 2)
    # BPTC(196,96): zero burst + 3 errors
    bptc = flips(N, 196, 3)
    # P25 half-rate trellis: 64 encoded paths tiled, noisy LLRs
    base = [H.p25_trellis_encode(rng)[1] for _ in range(64)]
    tx = np.stack([base[i % 64] for i in range(N)])
    bits = np.stack([(tx >> 1) & 1, tx & 1], axis=2).reshape(N, 196)
    llr = np.clip(np.rint(np.where(bits == 1, 200.0, -200.0) + rng.standard_normal((N, 196)) * 70.0), -32768, 32767).astype(np.int16)
    hard = ((llr[:, 0::2] > 0).astype(np.uint8) << 1) | (llr[:, 1::2] > 0).astype(np.uint8)
    # RS(36,20,17): zero word + 4 symbol errors
    rs = np.zeros((N, 36, 6), np.uint8)
    for _ in range(4):
        rs[np.arange(N), rng.integers(0, 36, N)] = rng.integers(0, 2, (N, 6))
    rs_par, rs_dat = np.ascontiguousarray(rs[:, :16].reshape(N, 96)), np.ascontiguousarray(rs[:, 16:].reshape(N, 120))
    # NID: zero codeword (NAC 0, DUID 0 = HDU, parity 0) + 4 errors, random reliabilities
    nid = flips(N, 63, 4)
    nid_rel = rng.integers(0, 256, (N, 63)).astype(np.uint8)
    nid_par = np.zeros(N, np.uint8)
    nid_prel = rng.integers(0, 256, N).astype(np.uint8)

    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_gol, d_ok = dev(gol), torch.zeros(N, dtype=torch.uint8, device="cuda")
    d_bptc, d_b96 = dev(bptc), torch.zeros((N, 96), dtype=torch.uint8, device="cuda")
    d_bR, d_berr = torch.zeros((N, 3), dtype=torch.uint8, device="cuda"), torch.zeros(N, dtype=torch.int32, device="cuda")
    d_llr, d_o12 = dev(llr), torch.zeros((N, 12), dtype=torch.uint8, device="cuda")
    d_met = torch.zeros(N, dtype=torch.int32, device="cuda")
    d_rsd, d_rsp, d_rst = dev(rs_dat), dev(rs_par), torch.zeros(N, dtype=torch.uint8, device="cuda")
    d_nid, d_nrel, d_npar, d_nprel = dev(nid), dev(nid_rel), dev(nid_par), dev(nid_prel)
    d_nst = torch.zeros(N, dtype=torch.int8, device="cuda")
    d_nnac = torch.zeros(N, dtype=torch.int32, device="cuda")
    d_nduid = torch.zeros(N, dtype=torch.uint8, device="cuda")
    d_nerr = torch.zeros(N, dtype=torch.int32, device="cuda")
    d_gol0, d_rsd0 = d_gol.clone(), d_rsd.clone()

    calls = {
        "golay_24_12": (lambda: (d_gol.copy_(d_gol0), lib.dsdneo_b200_fec_block_decode_batch(b200.FEC_GOLAY_24_12, d_gol.data_ptr(), None, d_ok.data_ptr(), N, None))[1],
                        48.0, "gq_decode_kernel"),
        "bptc_196x96": (lambda: lib.dsdneo_b200_bptc_196x96_batch(d_bptc.data_ptr(), 1, d_b96.data_ptr(), d_bR.data_ptr(), d_berr.data_ptr(), N, None),
                        299.0, "bptc_196x96_kernel"),
        "p25_12_soft_llr": (lambda: lib.dsdneo_b200_p25_12_soft_llr_batch(d_llr.data_ptr(), d_o12.data_ptr(), d_met.data_ptr(), N, None),
                            408.0, "p25_12_soft_llr_kernel"),
        "rs_36_20_17": (lambda: (d_rsd.copy_(d_rsd0), lib.dsdneo_b200_p25_rs_decode_batch(0, d_rsd.data_ptr(), d_rsp.data_ptr(), d_rst.data_ptr(), N, None))[1],
                        336.0, "p25_rs_decode_kernel"),
        "p25p1_nid_decode": (lambda: lib.dsdneo_b200_p25p1_nid_decode_batch(d_nid.data_ptr(), d_nrel.data_ptr(), None, d_npar.data_ptr(), d_nprel.data_ptr(), 64,
                                                                            d_nst.data_ptr(), d_nnac.data_ptr(), d_nduid.data_ptr(), d_nerr.data_ptr(), N, None),
                             136.0, "p25p1_nid_decode_kernel"),
    }
    for fn_, _, _ in calls.values():
        b200.check(fn_())
    torch.cuda.synchronize()
    b200.timing_enable(True)
    launches0 = b200.launch_count()
    steps = max(5, min(args.steps, 50))
    for _ in range(steps):
        for fn_, _, _ in calls.values():
            b200.check(fn_())
    torch.cuda.synchronize()
    rep = b200.timing_report()
    b200.timing_enable(False)
    launches = b200.launch_count() - launches0

    # reference functions, native loops on one core
    path = os.path.join(ROOT, "oracle", "_ref", "libdsdneo_ref_fast.so")
    cpu = {}
    if os.path.exists(path):
        L = C.CDLL(path)
        L.InitAllFecFunction()
        L.ref_fec_loop.restype = C.c_double
        L.ref_fec_loop.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_long)]
        packed = {
            "golay_24_12": (0, gol[:n_cpu]),
            "bptc_196x96": (1, bptc[:n_cpu]),
            "p25_12_soft_llr": (2, np.concatenate([hard[:n_cpu], llr[:n_cpu].view(np.uint8)], axis=1)),
            "rs_36_20_17": (3, np.concatenate([rs_dat[:n_cpu], rs_par[:n_cpu]], axis=1)),
            "p25p1_nid_decode": (4, np.concatenate([nid[:n_cpu], nid_rel[:n_cpu], nid_par[:n_cpu, None], nid_prel[:n_cpu, None]], axis=1)),
        }
        for name, (kind, arr) in packed.items():
            arr = np.ascontiguousarray(arr, np.uint8)
            chk = C.c_long(0)
            t1 = L.ref_fec_loop(kind, arr.ctypes.data, n_cpu, 1, C.byref(chk))
            reps = max(1, int(1.5 / max(t1, 1e-6)))
            t = L.ref_fec_loop(kind, arr.ctypes.data, n_cpu, reps, C.byref(chk))
            cpu[name] = n_cpu * reps / t
    out = {}
    for name, (_, bytes_per_item, kname) in calls.items():
        ms = rep[kname]["ms"] / rep[kname]["launches"]
        out[name] = {"kernel": kname, "gpu_ms_per_launch": ms, "gpu_items_per_s": N / (ms * 1e-3), "items_per_launch": N,
                     "algorithmic_bytes_per_item": bytes_per_item, "achieved_gbs": N * bytes_per_item / (ms * 1e-3) / 1e9,
                     "cpu_items_per_s_one_core": cpu.get(name)}
    peak, peak_src = measured_hbm_peak()
    print(json.dumps({
        "metric": "fec_items_per_s", "unit": "items/s", "n_gpus": 1, "steps": steps, "higher_is_better": True, "dtype": "u8",
        "data": "synthetic", "config": {"workload": "FEC leaves (section 8 rows a13-a16, f4; developer line): %d items per launch, "
                                        "device-resident inputs, reference functions natively timed on one host core" % N},
        "gpu_launches": int(launches), "hbm_peak_gbs": peak, "peak_source": peak_src, "kernels": out,
        "cpu_baseline": {"kind": "reference", "cores": 1, "unit": "items/s", "value": cpu,
                         "sample": "%d items per function, repeated for ~1.5 s each, perf-bench flags" % n_cpu}}))


def run_c4_workload(args):
    """Developer line for BASELINE config C4 (4096 DMR channels: channelizer + full_demod + 4FSK slicer + BPTC(196,96) / Golay +
    AMBE+2 frame ECC + MBE synthesis): runs tools/c4_bench.py and re-emits its result with the judged line's keys."""
    import subprocess

    n_ch = args.channels if args.channels != M else 4096
    steps = max(3, min(args.steps, 20))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "c4_bench.py"), str(n_ch), str(steps)], capture_output=True, text=True)
    if out.returncode != 0:
        raise SystemExit("bench.py --workload c4: tools/c4_bench.py failed:\n" + out.stderr[-2000:])
    r = json.loads(out.stdout.strip().splitlines()[-1])
    kern = r["kernels_ms"]
    dom = max(kern, key=kern.get)
    peak, peak_src = measured_hbm_peak()
    n_s = float(n_ch) * 49152
    algb = {"lpf_phase_kernel": 12.0 * n_s, "pfbn_kernel": 10.0 * n_s, "sps_fir_kernel": 8.0 * n_s, "disc_recurrence_kernel": 8.0 * n_s}
    roofline = None
    if dom in algb:
        a = algb[dom] / (kern[dom] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": algb[dom]}
    print(json.dumps({
        "metric": "iq_msps", "value": r["iq_msps"], "unit": "MS/s", "n_gpus": 1, "steps": steps, "warmup": 4, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "channels_at_realtime": r["channels_at_realtime"],
        "config": {"workload": "C4 (developer line): %d DMR base-station channels, 1.024 s per step: polyphase channelizer + full_demod on a "
                               "wideband cu8 tile, dmr matched filter + 4FSK slicer + BS DATA / VOICE sync hunt + burst cutters + BPTC(196,96) / "
                               "Golay(20,8) + AMBE+2 3600x2450 frame ECC on synthetic BS traffic at discriminator level; MBE synthesis timed "
                               "on the side (parity unpinned)" % n_ch,
                   "channels_per_gpu": n_ch, "pairs_per_channel_per_step": 49152},
        "roofline": roofline, "kernels": kern, "mbe_synth": r["mbe_synth"], "step_detail": r["decoded"], "cpu_baseline": None,
        "gpu_launches": None}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shard", default="bands", choices=["bands", "channels"],
                    help="bands (default): every GPU has its own channels, no collective; channels: ONE wideband stream, the raw IQ "
                         "tile reaches every GPU through one NCCL collective, each GPU channelizes and decodes its channel class")
    ap.add_argument("--channels", type=int, default=M)
    ap.add_argument("--channels-cu8", action="store_true",
                    help="(--shard channels) the channelizer hands the receive bank cu8 rows instead of cf32")
    ap.add_argument("--auto-reacquire", type=int, default=0,
                    help="(c3) developer switch: cfg.auto_reacquire_tiles of the receive bank (loss-of-sync watch on the device)")
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c4", "cqpsk", "fec"],
                    help="c3 (default, the judged line): 1024 P25 Phase 1 channels end to end; c2: 256-channel channelizer + "
                         "discriminator; cqpsk / fec: developer lines")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.workload == "c4":
        run_c4_workload(args)
    elif args.workload == "cqpsk":
        run_cqpsk_workload(args)
    elif args.workload == "fec":
        run_fec_workload(args)
    elif args.workload == "c2":
        if args.impl == "reference":
            run_c2_reference_arm(args)
        elif args.shard == "channels":
            run_channel_sharded(args)
        else:
            run_c2_b200_arm(args)
    elif args.impl == "reference":
        run_c3_reference_arm(args)
    elif args.shard == "channels":
        run_c3_one_stream(args)
    else:
        run_c3_b200_arm(args)


if __name__ == "__main__":
    main()
