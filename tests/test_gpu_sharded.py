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
