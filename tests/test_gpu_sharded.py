"""Channel-sharded multi-GPU mode on real GPUs: one NCCL broadcast of the raw IQ tile, every rank
demodulates its own channel range; the result must equal the unsharded front end bit for bit.  Needs >= 2 GPUs
(`gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_sharded_equals_unsharded_two_gpus(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import os, sys, json
        import numpy as np, torch, torch.distributed as dist
        sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
        import __graft_entry__ as g
        import _harness as H
        b200 = g.load_package()
        from dsdneo_b200 import shard
        rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        b200.init(local)
        M, bp, nb = 256, 1024, 2
        rng = np.random.default_rng(11)
        tiles = [H.synth_wideband(rng, M, bp * nb, [3, 100, 131, 250], snr_db=25.0)[0] for _ in range(3)]
        sf = shard.ShardedFrontend(b200, M, rank, world, 8, False, 12_288_000, bp)
        full = b200.Frontend(M, 8, False, 12_288_000, bp)
        ok = True
        bufs = [torch.from_numpy(t).cuda() if rank == 0 else torch.zeros((t.shape[0], 2), device="cuda") for t in tiles]
        for i, t in enumerate(tiles):
            if i + 1 < len(tiles):
                pass
            got = sf.process(bufs[i])
            if i + 1 < len(tiles):
                sf.prefetch(bufs[i + 1])          # next tile's broadcast overlaps nothing here, but exercises the path
            want = full.process(torch.from_numpy(t).cuda())[sf.lo:sf.hi]
            torch.cuda.synchronize()
            ok = ok and torch.equal(got.view(torch.int32), want.view(torch.int32))
        flags = [torch.zeros(1, device="cuda") for _ in range(world)]
        dist.all_gather(flags, torch.tensor([1.0 if ok else 0.0], device="cuda"))
        if rank == 0:
            print(json.dumps({"ok": [float(f) for f in flags], "ranges": [shard.channel_range(r, world, M) for r in range(world)]}))
        dist.destroy_process_group()
    """ % (ROOT, ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29541", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    import json
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["ok"] == [1.0, 1.0] and d["ranges"] == [[0, 128], [128, 256]]


# ---- ONE wideband stream -> bin-pruned channelizer -> P25 Phase 1 receive bank per rank (dsdneo_b200.shard.ShardedP25Rx) ----

def _one_stream_fixture(torch, b200, M, T, n_tiles, pairs):
    """Wideband cu8 stream (circular) carrying P25 C4FM traffic on every channel, built by the synthesis bank."""
    import numpy as np

    sys.path.insert(0, ROOT)
    import bench
    from dsdneo_b200 import shard

    base = bench.c3_base_iq(0)[:, :n_tiles * pairs]
    bc = torch.from_numpy((base.astype(np.float32) - 127.5) / 127.5).cuda()
    bc = torch.roll(torch.complex(bc[..., 0], bc[..., 1]), -(T - 1), dims=1).contiguous()
    idx = (torch.arange(M, device="cuda") * 7 + 3) % base.shape[0]
    proto = b200.Channelizer(M, T, True).prototype()
    return shard.synthesize_wideband(torch, bc, idx, M, proto, T), idx.cpu().numpy()


def test_one_stream_rank_emulation_on_one_gpu(gpu):
    """Two rank objects (world 2) on ONE GPU, fed the same tiles, against the unsharded chain (channelize_bins + bank called
    directly): identical records and dibits per channel class, every transmitted frame decoded, classes disjoint."""
    import torch

    sys.path.insert(0, ROOT)
    import bench
    from dsdneo_b200 import shard

    M, T, pairs, n_tiles, world = 512, 8, 16384, 6, 2
    wide, _ = _one_stream_fixture(torch, gpu, M, T, n_tiles, pairs)
    taps = bench._p25_filter_taps()
    tiles = [wide[t * pairs * M:(t + 1) * pairs * M] for t in range(n_tiles)]
    tail = wide[-(T - 1) * M:].contiguous()
    ranks = [shard.ShardedP25Rx(gpu, M, r, world, taps, pairs, block_pairs=8192) for r in range(world)]
    outs = [[sr.rx.alloc_device_out("cuda") for _ in range(n_tiles)] for sr in ranks]
    for sr in ranks:
        sr.cz.prime(tail)
    # pipelined use: tile i + 1 is distributed before tile i is submitted, nothing waits on the host in between
    for sr in ranks:
        sr.distribute(tiles[0], "broadcast")
    for t in range(n_tiles):
        for r, sr in enumerate(ranks):
            if t + 1 < n_tiles:
                sr.distribute(tiles[t + 1], "broadcast")
            tk = sr.submit(outs[r][t])
            if t == n_tiles - 1:
                sr.rx.wait(tk)
    torch.cuda.synchronize()
    total_ok = 0
    for r in range(world):
        cz = gpu.Channelizer(M, T, True)
        cz.prime(tail)
        rx = gpu.P25p1Rx(M // world, taps, block_pairs=8192, max_pairs_per_call=pairs, input_cu8=False)
        o = rx.alloc_device_out("cuda")
        for t in range(n_tiles):
            y = cz.channelize_bins(tiles[t], world, r)
            rx.process(y, pairs, o)
            fr, vo = rx.records(o)
            fs, vs = ranks[r].rx.records(outs[r][t])
            assert fr.tobytes() == fs.tobytes() and vo.tobytes() == vs.tobytes(), (r, t)
            assert torch.equal(o["counts"], outs[r][t]["counts"]) and torch.equal(o["dibits"], outs[r][t]["dibits"])
            if t >= 2:
                assert fr.size > 0 and (fr["nid_status"] > 0).all(), (r, t, int((fr["nid_status"] > 0).sum()), fr.size)
                total_ok += fr.size
    assert total_ok >= 4 * (M // 16)
    assert set(shard.channel_class(0, world, M)).isdisjoint(shard.channel_class(1, world, M))


def test_one_stream_two_gpus_nccl(tmp_path):
    """The same over NCCL on two GPUs (broadcast and all-gather forms of the collective): each rank's records equal the
    single-GPU computation of its channel class."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29500 + (os.getpid() % 400)
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(os.path.dirname(os.path.abspath(__file__)), "_one_stream_worker.py"), str(tmp_path)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists()
