"""Multi-GPU sharding of ONE wideband stream: channels are independent after the channelizer, so
GPU g owns the contiguous channel range [g*M/G, (g+1)*M/G) and the only collective is one broadcast of the raw IQ tile
from the ingest rank (NCCL over NVLink/NVSwitch on GPUs; any torch.distributed backend works for the host logic).

The reference has no counterpart (one process == one channel); this is the north star's
"channels shard trivially across the 8 GPUs of one box, so the only collective is a single NCCL broadcast of the raw
IQ tile".  PyTorch is plumbing here (process group, streams, device buffers); every kernel is behind the C-ABI.
"""
from __future__ import annotations

from typing import Optional, Tuple


def channel_range(rank: int, world: int, n_channels: int) -> Tuple[int, int]:
    """Contiguous, disjoint, exhaustive: the first n_channels % world ranks take one extra channel."""
    if not (0 <= rank < world) or n_channels < 0:
        raise ValueError("channel_range: bad rank/world")
    base, extra = divmod(n_channels, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_tile(tile, root: int = 0, group=None, async_op: bool = False):
    """The path's single collective: every rank ends up with the ingest rank's raw IQ tile (in place)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    return dist.broadcast(tile, src=root, group=group, async_op=async_op)


class ShardedFrontend:
    """Rank-local part of a channel-sharded front end: full channelizer, full_demod on this rank's channel range only.

    process(tile): `tile` is the wideband IQ tile on this rank's GPU ([n, 2] f32 or [n, 2] u8); on non-root ranks its
    contents are overwritten by the broadcast.  Returns this rank's [n_local, n / M] discriminator block.
    """

    def __init__(self, b200, n_channels: int, rank: int, world: int, taps_per_branch: int = 8, input_is_cu8: bool = False,
                 wideband_rate_hz: int = 12_288_000, block_pairs: int = 8192, root: int = 0, group=None,
                 profiles=None, squelch_levels=None):
        import torch

        self.b200, self.M, self.rank, self.world, self.root, self.group = b200, n_channels, rank, world, root, group
        self.lo, self.hi = channel_range(rank, world, n_channels)
        self.block_pairs = block_pairs
        self.cz = b200.Channelizer(n_channels, taps_per_branch, input_is_cu8)
        n_local = self.hi - self.lo
        self.bank = b200.DemodBank(
            n_local, wideband_rate_hz // n_channels, True,
            profiles=None if profiles is None else list(profiles[self.lo:self.hi]),
            squelch_levels=None if squelch_levels is None else list(squelch_levels[self.lo:self.hi]),
        ) if n_local > 0 else None
        self._chan = None
        self._bc_stream = torch.cuda.Stream() if world > 1 else None
        self._bc_done = None

    def close(self) -> None:
        self.cz.close()
        if self.bank is not None:
            self.bank.close()

    def prefetch(self, tile) -> None:
        """Start the broadcast of a later tile on a side stream so it overlaps the kernels of the current one."""
        import torch

        if self.world == 1:
            return
        self._bc_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._bc_stream):
            broadcast_tile(tile, self.root, self.group)
            self._bc_done = torch.cuda.Event()
            self._bc_done.record(self._bc_stream)
        self._prefetched = tile.data_ptr()

    def process(self, tile, out=None):
        import torch

        if self.world > 1:
            if getattr(self, "_prefetched", None) == tile.data_ptr() and self._bc_done is not None:
                torch.cuda.current_stream().wait_event(self._bc_done)
                self._prefetched = None
            else:
                broadcast_tile(tile, self.root, self.group)
        n_out = tile.shape[0] // self.M
        if self._chan is None or self._chan.shape[1] != n_out:
            self._chan = torch.empty((self.M, n_out, 2), device=tile.device, dtype=torch.float32)
        self.cz.channelize(tile, self._chan)
        if self.bank is None:
            return torch.empty((0, n_out), device=tile.device, dtype=torch.float32)
        return self.bank.full_demod(self._chan[self.lo:self.hi], self.block_pairs, n_out // self.block_pairs, out)
