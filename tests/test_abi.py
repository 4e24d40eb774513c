"""CPU tests of the boundary: the shared library loads without a GPU, exports every symbol the header declares,
fails loudly (no CPU fallback) when no device is present, and the reference-side glue compiles against the
reference's own headers."""
import os
import re
import subprocess

import pytest

import _harness as H

ROOT = H.ROOT
HEADER = os.path.join(ROOT, "include", "dsdneo_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dsdneo_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(b200):
    out = subprocess.run(["nm", "-D", "--defined-only", b200.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (dsdneo_b200_[a-z0-9_]+)", out))
    declared = _declared_symbols()
    assert len(declared) > 40
    missing = [s for s in declared if s not in exported]
    assert not missing, missing


def test_library_has_no_oracle_or_torch_dependency(b200):
    ldd = subprocess.run(["ldd", b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "oracle" not in ldd and "libdsdneo_ref" not in ldd
    sym = subprocess.run(["nm", "-D", b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in sym


def test_no_cpu_fallback_without_device(b200):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = b200.lib()
    assert L.dsdneo_b200_init(0) == b200.ENODEV
    assert b"no CPU fallback" in L.dsdneo_b200_last_error()
    with pytest.raises(b200.B200Error):
        b200.DemodBank(4)
    import numpy as np
    bits = np.zeros((2, 24), np.uint8)
    with pytest.raises(b200.B200Error):
        b200.fec_block_decode(b200.FEC_GOLAY_24_12, bits)


def test_host_side_lpf_design_matches_oracle(b200):
    import numpy as np

    for rate in (48000, 24000, 50000):
        for prof in range(6):
            assert np.array_equal(b200.channel_lpf_design(rate, prof).view(np.uint32), H.oracle_lpf_taps(rate, prof).view(np.uint32))
    with pytest.raises(b200.B200Error):
        b200.channel_lpf_design(96000, 4)  # 269 taps > 144: the reference would use its 63-tap fallback; unsupported here


@pytest.mark.skipif(not os.path.isdir("/root/reference/include"), reason="reference headers not present")
def test_reference_side_glue_compiles_against_reference_headers():
    R = "/root/reference"
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + R + "/include", "-I" + R + "/src"]
    subprocess.run(["g++", "-std=c++14", "-fsyntax-only"] + inc + [os.path.join(ROOT, "dsd-neo_b200/compat/full_demod_b200.cpp")], check=True)
    subprocess.run(["gcc", "-std=c11", "-fsyntax-only"] + inc + [os.path.join(ROOT, "dsd-neo_b200/compat/fec_b200.c")], check=True)
    # signatures of the shim must agree with the reference's own prototypes
    chk = ("#include <stdbool.h>\n#include <stdint.h>\n#include <dsd-neo/fec/block_codes.h>\n#include <dsd-neo/fec/bptc.h>\n"
           "#include <dsd-neo/protocol/p25/p25_12.h>\n#include \"%s\"\n" % os.path.join(ROOT, "dsd-neo_b200/compat/fec_b200.c"))
    subprocess.run(["gcc", "-std=c11", "-fsyntax-only", "-x", "c", "-"] + inc, input=chk, text=True, check=True)


def test_host_side_fll_band_edge_design_matches_oracle(b200):
    """dsdneo_b200_fll_band_edge_design (host/cqpsk_design.c) == the oracle's design (pinned to the reference's
    fll_band_edge_design_filter by tests/test_oracle_cqpsk.py), bit for bit, for every supported sps."""
    import numpy as np

    L, O = b200.lib(), H.oracle_cqpsk()
    for sps in range(2, 11):
        got = [np.zeros(48, np.float32) for _ in range(4)]
        want = [np.zeros(48, np.float32) for _ in range(4)]
        n = L.dsdneo_b200_fll_band_edge_design(sps, *[g.ctypes.data for g in got], 48)
        m = O.oracle_fll_band_edge_design(sps, *[H._ptr(w) for w in want], 48)
        assert n == m == 2 * sps + 1
        for g, w in zip(got, want):
            assert H.bits_equal(g, w)
    assert L.dsdneo_b200_cqpsk_block_capacity(2400, 5) >= 2400 // 5 + 2


def test_new_entry_points_fail_loudly_without_device(b200):
    """CQPSK bank / slicer, NID decode, soft word decoders and the frame cutters have no CPU fallback either."""
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(b200.B200Error):
        b200.CqpskBank(4, 24000)
    with pytest.raises(b200.B200Error):
        b200.CqpskSlicer(4)
    with pytest.raises(b200.B200Error):
        b200.p25p1_nid_decode(np.zeros((2, 63), np.uint8), None, None, np.zeros(2, np.uint8), None)
    L = b200.lib()
    bits, rel = np.zeros((2, 10), np.uint8), np.zeros((2, 10), np.int32)
    out, st = np.zeros((2, 10), np.uint8), np.zeros(2, np.uint8)
    assert L.dsdneo_b200_hamming_10_6_3_soft_batch_host(bits.ctypes.data, rel.ctypes.data, 1, 64, out.ctypes.data, st.ctypes.data, 2) == b200.ENODEV
    d, p, r = np.zeros((2, 6), np.uint8), np.zeros((2, 12), np.uint8), np.zeros((2, 18), np.int32)
    fx = np.zeros(2, np.int32)
    assert L.dsdneo_b200_p25_golay_soft_batch_host(b200.P25_WORD_GOLAY_24_6, d.ctypes.data, p.ctypes.data, r.ctypes.data, 1, 64,
                                                   st.ctypes.data, fx.ctypes.data, 2) == b200.ENODEV
    assert b"no CPU fallback" in L.dsdneo_b200_last_error()


def test_every_declared_entry_point_has_a_ctypes_prototype(b200):
    """The Python harness binds every entry point of include/dsdneo_b200.h with explicit argument types (ctypes' default int
    conversion would truncate 64-bit device pointers), and pointer-returning entry points with a pointer return type."""
    import ctypes as C

    L = b200.lib()
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    decl = re.findall(r"^([A-Za-z_][A-Za-z0-9_ \*]*?)\b(dsdneo_b200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.M | re.S)
    assert len(decl) > 150
    for ret, name, args in decl:
        fn = getattr(L, name)
        takes_args = args.strip() not in ("", "void")
        if takes_args:
            assert fn.argtypes is not None and len(fn.argtypes) == args.count(",") + 1, (name, fn.argtypes, args)
        if "*" in ret:
            assert fn.restype in (C.c_void_p, C.c_char_p), (name, fn.restype)
