"""torchrun worker of tests/test_gpu_sharded.py::test_one_stream_two_gpus_nccl (one process per GPU, NCCL)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import bench
from test_gpu_sharded import _one_stream_fixture

b200 = g.load_package()
from dsdneo_b200 import shard

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
b200.init(local)
M, T, pairs, n_tiles = 512, 8, 16384, 5
wide, _ = _one_stream_fixture(torch, b200, M, T, n_tiles, pairs)  # deterministic: the same on every rank
taps = bench._p25_filter_taps()
tiles = [wide[t * pairs * M:(t + 1) * pairs * M] for t in range(n_tiles)]
tail = wide[-(T - 1) * M:].contiguous()
n_slice = pairs * M // world
for mode in ("broadcast", "allgather"):
    sr = shard.ShardedP25Rx(b200, M, rank, world, taps, pairs, block_pairs=8192)
    sr.cz.prime(tail)
    outs = [sr.rx.alloc_device_out("cuda") for _ in range(n_tiles)]

    def src(t):
        if mode == "broadcast":
            return tiles[t] if rank == 0 else None  # only the ingest rank holds the stream
        return tiles[t][rank * n_slice:(rank + 1) * n_slice].cpu().pin_memory()  # every rank ingests its slice from the host

    sr.distribute(src(0), mode)
    for t in range(n_tiles):
        if t + 1 < n_tiles:
            sr.distribute(src(t + 1), mode)
        tk = sr.submit(outs[t])
    sr.rx.wait(tk)
    torch.cuda.synchronize()
    cz = b200.Channelizer(M, T, True)
    cz.prime(tail)
    rx = b200.P25p1Rx(M // world, taps, block_pairs=8192, max_pairs_per_call=pairs, input_cu8=False)
    o = rx.alloc_device_out("cuda")
    n_ok = 0
    for t in range(n_tiles):
        rx.process(cz.channelize_bins(tiles[t], world, rank), pairs, o)
        fr, vo = rx.records(o)
        fs, vs = sr.rx.records(outs[t])
        assert fr.tobytes() == fs.tobytes() and vo.tobytes() == vs.tobytes(), (mode, rank, t)
        assert torch.equal(o["dibits"], outs[t]["dibits"]) and torch.equal(o["counts"], outs[t]["counts"])
        if t >= 2:
            assert fr.size > 0 and (fr["nid_status"] > 0).all(), (mode, rank, t)
            n_ok += fr.size
    assert n_ok > 0
    sr.close()
dist.barrier()
open(os.path.join(sys.argv[1], "rank%d.ok" % rank), "w").write("ok")
dist.destroy_process_group()
