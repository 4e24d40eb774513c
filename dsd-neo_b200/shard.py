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


# --------------------------------------------------------------------------------------------------------------------
# One wideband stream -> P25 Phase 1 receive banks on every GPU (BASELINE configs C4 / C5 shape)
# --------------------------------------------------------------------------------------------------------------------

def channel_class(rank: int, world: int, n_channels: int):
    """Channels owned by `rank` under the bin-pruned channelizer: k = rank (mod world).  Row k' of the rank's channelizer
    output (and of its receive bank) is channel world * k' + rank; disjoint and exhaustive over ranks."""
    if not (0 <= rank < world) or n_channels % world:
        raise ValueError("channel_class: world must divide n_channels")
    return range(rank, n_channels, world)


def synthesize_wideband(torch, chan_c64, chan_of_bin, n_channels: int, prototype, taps_per_branch: int, rows_per_block: int = 2048,
                        rms: float = 0.25, seed: int = 1):
    """Test / bench signal source (not on the product path): a polyphase SYNTHESIS bank, the transpose of the channelizer.

    chan_c64: [n_base, n] complex64 on the GPU, the channel-rate signals; chan_of_bin: int64 [M] on the GPU, which base
    signal channel k carries (-1: empty).  Every channel gets a random constant phase so the sum is noise-like.  Output:
    cu8 [n * M, 2] (on the GPU) on a CIRCULAR time axis:
        x[n M + r] = sum_q M h[q M + r] V[n - q][r],   V[n][r] = sum_k c_k[n] exp(+j 2 pi k r / M)
    i.e. every c_k interpolated by M with the channelizer's own prototype and moved to k * fs / M."""
    M, T = n_channels, taps_per_branch
    dev = chan_c64.device
    n = chan_c64.shape[1]
    g = torch.Generator(device="cpu").manual_seed(seed)
    theta = (torch.rand(M, generator=g) * 6.283185307179586).to(dev)
    rot = torch.polar(torch.ones(M, device=dev), theta) * (chan_of_bin >= 0)
    idx = chan_of_bin.clamp(min=0)
    hs = (torch.as_tensor(prototype, dtype=torch.float32, device=dev) * M).reshape(T, M)  # hs[q][r] = M h[q M + r]
    amp = float(chan_c64.abs().pow(2).mean().sqrt().item()) * float((chan_of_bin >= 0).sum().item()) ** 0.5
    scale = rms / max(amp, 1e-12)
    out = torch.empty((n * M, 2), dtype=torch.uint8, device=dev)
    ov = out.view(n, M, 2)
    for a in range(0, n, rows_per_block):
        b = min(n, a + rows_per_block)
        rows = (torch.arange(a - (T - 1), b, device=dev) % n)
        c = chan_c64[:, rows].t()[:, idx] * rot  # [rows, M]
        V = torch.fft.ifft(c, dim=1) * M
        x = torch.zeros((b - a, M), dtype=torch.complex64, device=dev)
        for q in range(T):
            x += hs[q] * V[T - 1 - q:T - 1 - q + (b - a)]
        x *= scale
        ov[a:b, :, 0] = torch.clamp(torch.round(x.real * 127.5 + 127.5), 0, 255).to(torch.uint8)
        ov[a:b, :, 1] = torch.clamp(torch.round(x.imag * 127.5 + 127.5), 0, 255).to(torch.uint8)
    return out


class ShardedP25Rx:
    """Rank-local part of ONE wideband stream received on `world` GPUs: the raw cu8 tile reaches every GPU through the
    path's single collective (a broadcast from the ingest rank, or an all-gather of per-rank slices when every rank
    ingests 1/world of the tile from its own host), the bin-pruned channelizer keeps channels k = rank (mod world), and
    the P25 Phase 1 receive bank (dsdneo_b200_p25p1_rx_*) decodes them.  Double-buffered: the collective of tile i + 1
    runs on a side stream under the kernels of tile i."""

    def __init__(self, b200, n_channels: int, rank: int, world: int, p25_taps, pairs_per_tile: int, rate_hz: int = 48000,
                 block_pairs: int = 8192, taps_per_branch: int = 8, root: int = 0, group=None, max_hits: int = 32, device=None,
                 acquire_tiles: int = 0, auto_reacquire_tiles: int = 0, channels_cu8: bool = False, channel_gain: Optional[float] = None):
        import torch

        self.b200, self.M, self.rank, self.world, self.root, self.group = b200, n_channels, rank, world, root, group
        self.n_local = len(channel_class(rank, world, n_channels))
        self.pairs = pairs_per_tile
        self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.cz = b200.Channelizer(n_channels, taps_per_branch, input_is_cu8=True)
        # channels_cu8: the channelizer hands the bank cu8 rows (its native capture format, 2 B per sample through HBM instead of
        # 8), scaled by channel_gain (default sqrt(M): a channel of an evenly loaded band back at the wideband level)
        self.channels_cu8 = channels_cu8
        self.channel_gain = float(channel_gain) if channel_gain is not None else float(n_channels) ** 0.5
        self.rx = b200.P25p1Rx(self.n_local, p25_taps, rate_hz=rate_hz, block_pairs=block_pairs, max_pairs_per_call=pairs_per_tile,
                               input_cu8=channels_cu8, max_hits=max_hits, acquire_tiles=acquire_tiles,
                               auto_reacquire_tiles=auto_reacquire_tiles)
        self.chan = [torch.empty((self.n_local, pairs_per_tile, 2), dtype=torch.uint8 if channels_cu8 else torch.float32, device=self.dev)
                     for _ in range(2)]
        self.raw = [torch.empty((pairs_per_tile * n_channels, 2), dtype=torch.uint8, device=self.dev) for _ in range(2)]
        self._side = torch.cuda.Stream(device=self.dev)
        self._arrived = [None, None]    # event: raw[b] holds its tile
        self._raw_free = [None, None]   # event: the channelizer has read raw[b]
        self._tickets = [None, None]    # bank ticket that reads chan[b]
        self._n_sub = 0                 # tiles submitted
        self._n_dist = 0                # tiles whose collective was issued

    def close(self):
        self.cz.close()
        self.rx.close()

    def distribute(self, src, mode: str = "broadcast"):
        """Issue the collective of the next tile on the side stream.  mode "broadcast": `src` is the whole tile on the root
        (ignored elsewhere); "allgather": `src` is this rank's slice [pairs / world * M, 2] (device or pinned host)."""
        import torch
        import torch.distributed as dist

        b = self._n_dist & 1
        self._n_dist += 1
        side = self._side
        side.wait_stream(torch.cuda.current_stream(self.dev))
        if self._raw_free[b] is not None:
            side.wait_event(self._raw_free[b])
        with torch.cuda.stream(side):
            if mode == "broadcast":
                if self.rank == self.root:
                    self.raw[b].copy_(src, non_blocking=True)
                elif not dist.is_initialized():
                    self.raw[b].copy_(src, non_blocking=True)  # single-process emulation of several ranks (tests)
                if self.world > 1 and dist.is_initialized():
                    dist.broadcast(self.raw[b], src=self.root, group=self.group)
            else:
                n_slice = self.raw[b].shape[0] // self.world
                mine = self.raw[b][self.rank * n_slice:(self.rank + 1) * n_slice]
                mine.copy_(src, non_blocking=True)
                if self.world > 1:
                    dist.all_gather_into_tensor(self.raw[b], mine.clone() if False else mine, group=self.group)
            ev = torch.cuda.Event()
            ev.record(side)
        self._arrived[b] = ev

    def submit(self, out, stream=None):
        """Channelize + queue the oldest distributed tile on the receive bank; returns the bank's ticket."""
        import torch

        if stream is None:
            stream = torch.cuda.current_stream(self.dev)
        assert self._n_sub < self._n_dist, "submit() without a distributed tile"
        b = self._n_sub & 1
        self._n_sub += 1
        stream.wait_event(self._arrived[b])
        if self._tickets[b] is not None:
            self.rx.input_consumed(self._tickets[b], stream)  # the bank's first stage has read chan[b]
        if self.channels_cu8:
            self.cz.channelize_bins_cu8(self.raw[b], self.world, self.rank, self.channel_gain, self.chan[b], stream=stream)
        else:
            self.cz.channelize_bins(self.raw[b], self.world, self.rank, self.chan[b], stream=stream)
        ev = torch.cuda.Event()
        ev.record(stream)
        self._raw_free[b] = ev
        self._tickets[b] = self.rx.submit(self.chan[b], self.pairs, out, stream)
        return self._tickets[b]
