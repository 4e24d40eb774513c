"""N>1 host logic on CPU: two gloo ranks shard bands exactly like bench.py does (one independent wideband band per rank,
no data-path collective), reduce their step times with MAX, and rank 0 aggregates the whole-job rate."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_band_sharding_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import os, sys, json
        import torch, torch.distributed as dist
        sys.path.insert(0, %r)
        import bench
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        # each rank synthesises ITS band (seeded by rank) -- reduced size so the CPU test stays fast
        bench.N_IN = 256 * 40
        x = bench.make_wideband(torch, torch.device("cpu"), seed=rank)
        sig = torch.tensor([float(x.double().abs().sum())])
        gathered = [torch.zeros(1) for _ in range(world)]
        dist.all_gather(gathered, sig)
        t = torch.tensor([1.0 + rank], dtype=torch.float64)       # pretend step time: rank 1 is slower
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            value = world * bench.N_IN / (float(t) * 1e-3) / 1e6
            print(json.dumps({"max_ms": float(t), "bands_differ": float(gathered[0]) != float(gathered[1]), "value": value}))
        dist.destroy_process_group()
    """ % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["max_ms"] == 2.0 and d["bands_differ"] is True
    assert abs(d["value"] - 2 * 256 * 40 / 2e-3 / 1e6) < 1e-9


def test_reference_arm_runs_on_rank0_only(tmp_path):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_channel_ranges_partition():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    shard = __import__("importlib").import_module("dsdneo_b200.shard") if g.load_package() else None
    for n_ch in (256, 1024, 4096, 8192, 250, 7):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard.channel_range(r, world, n_ch) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n_ch
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_tile_broadcast_gloo(tmp_path):
    """The channel-sharded mode's one collective on CPU/gloo: the ingest rank's raw IQ tile reaches rank 1 unchanged
    (in place), the two ranks' channel ranges tile [0, M), and the gathered per-rank results come back in channel order."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import os, sys, json
        import torch, torch.distributed as dist
        sys.path.insert(0, %r)
        import __graft_entry__ as g
        g.load_package()
        from dsdneo_b200 import shard
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        M, n = 256, 256 * 64
        gen = torch.Generator().manual_seed(5)
        ref_tile = torch.randint(0, 256, (n, 2), dtype=torch.uint8, generator=gen)     # cu8 IQ tile, same seed everywhere
        tile = ref_tile.clone() if rank == 0 else torch.zeros_like(ref_tile)
        shard.broadcast_tile(tile, root=0)
        lo, hi = shard.channel_range(rank, world, M)
        # stand-in for the per-rank kernels: each rank labels its own channels; rank 0 gathers in channel order
        mine = torch.arange(lo, hi, dtype=torch.int64)
        parts = [torch.zeros(shard.channel_range(r, world, M)[1] - shard.channel_range(r, world, M)[0], dtype=torch.int64) for r in range(world)]
        dist.all_gather(parts, mine)
        ok = torch.equal(tile, ref_tile)
        oks = [torch.zeros(1) for _ in range(world)]
        dist.all_gather(oks, torch.tensor([1.0 if ok else 0.0]))
        if rank == 0:
            print(json.dumps({"tile_ok": [float(o) for o in oks], "order_ok": torch.equal(torch.cat(parts), torch.arange(M))}))
        dist.destroy_process_group()
    """ % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29534", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d == {"tile_ok": [1.0, 1.0], "order_ok": True}
