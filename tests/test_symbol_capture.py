"""The reference's DSDNSYM2 symbol-capture format as harness I/O: byte-identical to what the
unmodified reference writer produces, round trip, header rules of the reference reader."""
import ctypes as C
import os

import numpy as np
import pytest

import _harness as H

needs_ref = pytest.mark.skipif(not H.ref_available("par"), reason="oracle/_ref not built (no /root/reference)")


def _b200():
    import __graft_entry__ as g

    return g.load_package()


def _records(rng, n):
    d = rng.integers(0, 4, n).astype(np.uint8)
    r = rng.integers(0, 256, n).astype(np.uint8)
    l = rng.integers(-32768, 32768, (n, 2)).astype(np.int16)
    s = (rng.standard_normal(n) * 20000).astype(np.float32)
    s[:3] = [0.0, -0.0, np.float32(3.4e38)]
    return d, r, l, s


@needs_ref
def test_pack_equals_the_reference_writer(tmp_path):
    """dsdneo_b200_symbol_capture_pack == header of openSymbolOutFile + write_symbol_capture_record of the unmodified
    reference (driven through oracle/ref_shim_symbol.c), byte for byte."""
    b = _b200()
    R = C.CDLL(H._ref_path("par"))
    rng = np.random.default_rng(1)
    d, r, l, s = _records(rng, 777)
    path = str(tmp_path / "ref.bin")
    n = R.ref_symbol_capture_write(path.encode(), H._ptr(d, H.u8p), H._ptr(r, H.u8p), l.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(s), d.size)
    assert n == d.size
    want = open(path, "rb").read()
    assert b.symbol_capture_pack(d, r, l, s) == want
    mine = str(tmp_path / "mine.bin")
    L = b.lib()
    b.check(L.dsdneo_b200_symbol_capture_write_file(mine.encode(), 0, d.ctypes.data, r.ctypes.data, l.ctypes.data, s.ctypes.data, 400))
    b.check(L.dsdneo_b200_symbol_capture_write_file(mine.encode(), 1, d[400:].ctypes.data, r[400:].ctypes.data, l[400:].ctypes.data,
                                                    s[400:].ctypes.data, d.size - 400))
    assert open(mine, "rb").read() == want


def test_golden_bytes_and_round_trip():
    """Layout pinned without the reference tree: header 'DSDNSYM2',2,10,0..; record = dibit, reliability, two LE int16, LE f32."""
    b = _b200()
    blob = b.symbol_capture_pack([1, 3], [200, 7], [[-2, 300], [32767, -32768]], np.array([1.0, -2.5], np.float32))
    assert blob == (b"DSDNSYM2\x02\x0a\x00\x00\x00\x00\x00\x00"
                    b"\x01\xc8\xfe\xff\x2c\x01\x00\x00\x80\x3f"
                    b"\x03\x07\xff\x7f\x00\x80\x00\x00\x20\xc0")
    rng = np.random.default_rng(2)
    d, r, l, s = _records(rng, 1000)
    for hdr in (True, False):
        dd, rr, ll, ss = b.symbol_capture_unpack(b.symbol_capture_pack(d, r, l, s, with_header=hdr))
        assert np.array_equal(dd, d) and np.array_equal(rr, r) and np.array_equal(ll, l)
        assert np.array_equal(ss.view(np.uint32), s.view(np.uint32))
    with pytest.raises(b.B200Error):  # the reference reader rejects a soft header with another version / record size
        b.symbol_capture_unpack(b"DSDNSYM2\x03\x0a" + bytes(6) + bytes(10))


@pytest.mark.gpu
def test_symbolizer_output_to_capture_file(gpu, tmp_path):
    """A channel demodulated on the GPU, written as a capture file: the records carry exactly the symbolizer's outputs."""
    import torch
    from test_gpu_symbolizer import _taps

    rng = np.random.default_rng(3)
    dib = rng.integers(0, 4, 600)
    x = H.synth_c4fm_disc(rng, dib, 9000.0, 400.0)
    sy = gpu.Symbolizer(1, 48000, 4800, filters=_taps())
    sy.set_class([gpu.sym_class_from_synctype(0, 0)])
    res = sy.run(torch.from_numpy(x[None, :]).cuda(), x.size)
    torch.cuda.synchronize()
    k = int(res["count"][0])
    blob = gpu.symbol_capture_pack(res["dibits"][0, :k].cpu().numpy(), res["reliability"][0, :k].cpu().numpy(),
                                   res["llr"][0, :k].cpu().numpy(), res["symbols"][0, :k].cpu().numpy())
    assert len(blob) == 16 + 10 * k
    d, r, l, s = gpu.symbol_capture_unpack(blob)
    assert np.array_equal(d, res["dibits"][0, :k].cpu().numpy()) and np.array_equal(l, res["llr"][0, :k].cpu().numpy())
