"""Row-band execution on several GPUs: one process per GPU, `torch.distributed` for the
plumbing (NCCL over NVLink on the GPUs, gloo in the CPU tests), the stepping itself in the CUDA library.

The reference is single-device; the decomposition follows from its stencil radius of exactly one cell in
both phases (steppers.fut:19,33-34,66,77-79).  Per tick each band
  1. swaps two heading/preference rows and one colour row with each neighbour        (state message)
  2. runs k_u_local, swaps its boundary row of calc / my|mx (so the ghost rows hold the neighbours' rows after
     their block-local sweeps) and runs the frontier iterations (k_u_iter)
  3. swaps its boundary row of calc / my|mx again, ORs the neighbour's row into its ghost row and wakes the owned
     words next to every ghost word that changed; all bands all-reduce (MAX) the number of woken words and
     resume k_u_iter until nobody was woken -- a waiting chain that crosses a cut costs one extra round
  4. [perfect mode] resolves its rings (a ring that straddles a cut is judged by every band it touches, the verdict
     bits are AND-ed over the bands, then it is marked) and swaps the boundary my|mx once more
  5. moves (k_move)
Messages are a few hundred KB, so this is latency-, not bandwidth-bound.

Transports.  BandRunner.tick is the protocol spelled out phase by phase over torch.distributed (what the gloo CPU tests
drive, with bands backed by the CPU emulation of the kernel programs).  On GPUs the ticks run natively inside the library
(flowamok_band_step): over the P2P transport (DeviceBand.p2p_init: peer-mapped mailboxes over NVLink, remote stores +
device-side flags, the convergence rounds and their termination vote inside one persistent kernel -- no NCCL call and no
host synchronisation per tick), or, for comparison, over NCCL send/recv with one host round trip per tick (comm_init).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist


def split_rows(gh: int, world: int, align: int = 16):
    """Row ranges of the bands: as even as possible, boundaries on multiples of `align`."""
    units = (gh + align - 1) // align
    cuts = [min(gh, align * ((units * r) // world)) for r in range(world + 1)]
    cuts[-1] = gh
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class DeviceBand:
    """A band living in the CUDA library (flowamok_band_* of include/flowamok.h)."""

    def __init__(self, ctx, gh_global, gw, row_begin, row_end):
        from . import Grid
        self.ctx, self.L = ctx, ctx.L
        self.gh_global, self.gw, self.rows = gh_global, gw, (row_begin, row_end)
        h = C.c_void_p()
        ctx.check(self.L.flowamok_band_new(ctx.h, C.byref(h), gh_global, gw, row_begin, row_end))
        self.grid = Grid.__new__(Grid)      # reuse upload/download/chooser/cycles/synth of the Grid wrapper
        self.grid.ctx, self.grid.L, self.grid.gh, self.grid.gw, self.grid.h, self.grid.joined_rng = \
            ctx, ctx.L, gh_global, gw, h, None
        sb, ub = C.c_int64(), C.c_int64()
        ctx.check(self.L.flowamok_band_msg_bytes(ctx.h, h, C.byref(sb), C.byref(ub)))
        self.state_bytes, self.u_bytes = sb.value, ub.value
        self.device = torch.device("cuda", ctx.device)
        # Messages and the woken counter are torch tensors: torch (NCCL, .item()) touches them on ITS current
        # stream.  If the library context was not created on that stream, fence at every hand-over.
        self.foreign_stream = ctx.stream_ptr != torch.cuda.current_stream(self.device).cuda_stream

    def _fence(self):
        if self.foreign_stream:
            torch.cuda.synchronize(self.device)

    native = False

    def set_cuts(self, cuts):
        """row boundaries of ALL bands (world + 1 numbers): needed before cycles that may straddle a cut are installed"""
        a = (C.c_int64 * len(cuts))(*cuts)
        self._c(self.L.flowamok_band_set_cuts(self.ctx.h, self.grid.h, len(cuts) - 1, a))

    def p2p_open(self, rank, world):
        """allocate + export this band's mailbox: (64-byte CUDA IPC handle, raw device pointer)"""
        h = (C.c_ubyte * 64)()
        ptr = C.c_void_p()
        self._c(self.L.flowamok_band_p2p_open(self.ctx.h, self.grid.h, rank, world, h, C.byref(ptr)))
        return bytes(h), ptr.value

    def p2p_connect(self, handles=None, ptrs=None, bands_on_device=1):
        world = len(handles) if handles is not None else len(ptrs)
        hb = b"".join(handles) if handles is not None else None
        pa = (C.c_void_p * world)(*[p or None for p in ptrs]) if ptrs is not None else None
        self._c(self.L.flowamok_band_p2p_connect(self.ctx.h, self.grid.h, hb, pa, bands_on_device))
        self.native = True

    def p2p_init(self, rank, world, group=None):
        """One band per process / GPU: exchange the mailbox handles through torch.distributed and map the peers.
        flowamok_band_set_cuts and the cycles must be in place."""
        handle, _ = self.p2p_open(rank, world)
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        self.p2p_connect(handles=handles)

    def step_profile(self, perfect, n_ticks=20):
        ms = (C.c_double * 5)()
        self._c(self.L.flowamok_band_step_profile(self.ctx.h, self.grid.h, int(perfect), n_ticks, ms))
        return dict(k_band_xstate=ms[0], k_u_local=ms[1], k_rings=ms[2], k_u_iter=ms[3], k_move=ms[4])

    def p2p_stats(self):
        r, w = C.c_int64(), C.c_int64()
        self._c(self.L.flowamok_band_p2p_stats(self.ctx.h, self.grid.h, C.byref(r), C.byref(w)))
        return dict(rounds=r.value, wait_ns=w.value)

    def p2p_close(self):
        """unmap the peers' mailboxes; synchronise all processes before freeing any band afterwards"""
        self._c(self.L.flowamok_band_p2p_close(self.ctx.h, self.grid.h))
        self.native = False

    def ring_words(self):
        n = C.c_int64()
        self._c(self.L.flowamok_band_ring_words(self.ctx.h, self.grid.h, C.byref(n)))
        return n.value

    def ring_buffer(self):
        return torch.zeros(max(1, self.ring_words()), dtype=torch.int32, device=self.device)

    def ring_verdicts(self, buf):
        self._c(self.L.flowamok_band_ring_verdicts(self.ctx.h, self.grid.h, buf.data_ptr()))
        self._fence()

    def rings_apply(self, buf):
        self._fence()
        self._c(self.L.flowamok_band_rings_apply(self.ctx.h, self.grid.h, buf.data_ptr()))

    def comm_init(self, rank, world, group=None):
        """Join the library's own NCCL communicator (unique id from rank 0, broadcast through torch.distributed);
        afterwards BandRunner.step runs whole ticks inside the library (flowamok_band_step)."""
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            if self.L.flowamok_band_comm_id(ident):
                raise RuntimeError("NCCL is not available to the library")
        obj = [bytes(ident)]
        dist.broadcast_object_list(obj, src=0, group=group)
        self._c(self.L.flowamok_band_comm_init(self.ctx.h, obj[0], rank, world))
        self.native = True

    def step_native(self, perfect, n_ticks):
        r = C.c_int64()
        self._c(self.L.flowamok_band_step(self.ctx.h, self.grid.h, int(perfect), n_ticks, C.byref(r)))
        return r.value

    def upload_cars(self, dy, dx, color, pref, rng_state):
        """whole-grid base pointers (ints); only the owned rows are read"""
        self._c(self.L.flowamok_grid_upload_cars(self.ctx.h, self.grid.h, dy, dx, color, pref, rng_state))

    def buffer(self, nbytes):
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def counter(self):
        return torch.zeros(1, dtype=torch.int32, device=self.device)

    def _c(self, rc):
        self.ctx.check(rc)

    def pack_state(self, side, buf):
        self._c(self.L.flowamok_band_pack_state(self.ctx.h, self.grid.h, side, buf.data_ptr()))
        self._fence()

    def unpack_state(self, side, buf):
        self._fence()
        self._c(self.L.flowamok_band_unpack_state(self.ctx.h, self.grid.h, side, buf.data_ptr()))

    def u_init(self):
        self._c(self.L.flowamok_band_u_init(self.ctx.h, self.grid.h))

    def u_iter(self):
        self._c(self.L.flowamok_band_u_iter(self.ctx.h, self.grid.h))

    def pack_u(self, side, buf):
        self._c(self.L.flowamok_band_pack_u(self.ctx.h, self.grid.h, side, buf.data_ptr()))
        self._fence()

    def merge_u(self, side, buf, wake):
        self._fence()
        self._c(self.L.flowamok_band_merge_u(self.ctx.h, self.grid.h, side, buf.data_ptr(), int(wake), None))

    def woken(self, counter):
        self._c(self.L.flowamok_band_woken(self.ctx.h, self.grid.h, counter.data_ptr()))
        self._fence()

    def u_resume(self):
        self._c(self.L.flowamok_band_u_resume(self.ctx.h, self.grid.h))

    def rings(self):
        self._c(self.L.flowamok_band_rings(self.ctx.h, self.grid.h))

    def move(self):
        self._c(self.L.flowamok_band_move(self.ctx.h, self.grid.h))

    def leaks(self):
        n = C.c_int64()
        self._c(self.L.flowamok_band_leaks(self.ctx.h, self.grid.h, C.byref(n)))
        return n.value


class BandRunner:
    """Drives one band; `band` is a DeviceBand (or the CPU emulation used by tests/test_bands_cpu.py)."""

    def __init__(self, band, rank: int, world: int, group=None):
        self.b, self.rank, self.world, self.group = band, rank, world, group
        self.up = rank - 1 if rank > 0 else None          # neighbour with smaller rows (side 0)
        self.down = rank + 1 if rank < world - 1 else None
        mk = band.buffer
        self.s_send = [mk(band.state_bytes), mk(band.state_bytes)]
        self.s_recv = [mk(band.state_bytes), mk(band.state_bytes)]
        self.u_send = [mk(band.u_bytes), mk(band.u_bytes)]
        self.u_recv = [mk(band.u_bytes), mk(band.u_bytes)]
        self.cnt = band.counter()
        self.ring_buf = None
        self.rounds = 0
        self.ticks = 0

    def _neigh(self):
        return [(0, self.up), (1, self.down)]

    def _swap(self, send, recv):
        ops = []
        for side, peer in self._neigh():
            if peer is None:
                continue
            ops.append(dist.P2POp(dist.isend, send[side], peer, self.group))
            ops.append(dist.P2POp(dist.irecv, recv[side], peer, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def exchange_state(self):
        for side, peer in self._neigh():
            if peer is not None:
                self.b.pack_state(side, self.s_send[side])
        self._swap(self.s_send, self.s_recv)
        for side, peer in self._neigh():
            if peer is not None:
                self.b.unpack_state(side, self.s_recv[side])

    def exchange_u(self, wake: bool):
        for side, peer in self._neigh():
            if peer is not None:
                self.b.pack_u(side, self.u_send[side])
        self._swap(self.u_send, self.u_recv)
        for side, peer in self._neigh():
            if peer is not None:
                self.b.merge_u(side, self.u_recv[side], wake)

    def tick(self, perfect: bool):
        if self.world > 1:
            self.exchange_state()
        self.b.u_init()
        if self.world > 1:
            self.exchange_u(wake=False)      # ghost rows := the neighbours' rows after their block-local sweeps
        self.b.u_iter()
        if self.world > 1:
            while True:
                self.exchange_u(wake=True)
                self.b.woken(self.cnt)
                dist.all_reduce(self.cnt, op=dist.ReduceOp.MAX, group=self.group)
                self.rounds += 1
                if int(self.cnt.item()) == 0:        # the one host round trip of the tick
                    break
                self.b.u_resume()
        if perfect:
            self.b.rings()
            if self.world > 1:
                if self.b.ring_words() > 0:           # rings that straddle a cut: AND the bands' verdicts, then mark
                    if self.ring_buf is None:
                        self.ring_buf = self.b.ring_buffer()
                    self.b.ring_verdicts(self.ring_buf)
                    dist.all_reduce(self.ring_buf, op=dist.ReduceOp.BAND, group=self.group)
                    self.b.rings_apply(self.ring_buf)
                self.exchange_u(wake=False)
        self.b.move()
        self.ticks += 1

    def step(self, perfect: bool, n_ticks: int):
        if getattr(self.b, "native", False):      # the same protocol, driven inside the library over its own NCCL comm
            self.rounds = self.b.step_native(perfect, n_ticks)
            self.ticks += n_ticks
            return
        for _ in range(n_ticks):
            self.tick(perfect)


class LocalBands:
    """All bands in ONE process on one device (no torch.distributed): the same protocol with the messages handed
    over directly.  Used to validate the band entry points on a single GPU and for debugging."""

    def __init__(self, bands):
        self.bands = bands
        n = len(bands)
        self.s = [[b.buffer(b.state_bytes), b.buffer(b.state_bytes)] for b in bands]
        self.u = [[b.buffer(b.u_bytes), b.buffer(b.u_bytes)] for b in bands]
        self.cnt = [b.counter() for b in bands]
        self.n = n
        self.rounds = 0

    def _pairs(self):
        # (sender, side it sends to, receiver, side the receiver files it under)
        for r in range(self.n):
            if r > 0:
                yield r, 0, r - 1, 1
            if r < self.n - 1:
                yield r, 1, r + 1, 0

    def tick(self, perfect):
        B = self.bands
        if self.n > 1:
            for r, side, _q, _qs in self._pairs():
                B[r].pack_state(side, self.s[r][side])
            for r, side, q, qs in self._pairs():
                B[q].unpack_state(qs, self.s[r][side])
        for b in B:
            b.u_init()
        if self.n > 1:
            for r, side, _q, _qs in self._pairs():
                B[r].pack_u(side, self.u[r][side])
            for r, side, q, qs in self._pairs():
                B[q].merge_u(qs, self.u[r][side], False)
        for b in B:
            b.u_iter()
        while self.n > 1:
            for r, side, _q, _qs in self._pairs():
                B[r].pack_u(side, self.u[r][side])
            for r, side, q, qs in self._pairs():
                B[q].merge_u(qs, self.u[r][side], True)
            woken = 0
            for r, b in enumerate(B):
                b.woken(self.cnt[r])
                woken = max(woken, int(self.cnt[r].item()))
            self.rounds += 1
            if woken == 0:
                break
            for b in B:
                b.u_resume()
        if perfect:
            for b in B:
                b.rings()
            if self.n > 1 and B[0].ring_words() > 0:
                bufs = [b.ring_buffer() for b in B]
                for b, v in zip(B, bufs):
                    b.ring_verdicts(v)
                for v in bufs[1:]:
                    bufs[0] &= v
                for b in B:
                    b.rings_apply(bufs[0])
            if self.n > 1:
                for r, side, _q, _qs in self._pairs():
                    B[r].pack_u(side, self.u[r][side])
                for r, side, q, qs in self._pairs():
                    B[q].merge_u(qs, self.u[r][side], False)
        for b in B:
            b.move()

    def step(self, perfect, n_ticks):
        for _ in range(n_ticks):
            self.tick(perfect)


class LocalBandsP2P:
    """All bands in ONE process on one device, each with its own context (= its own stream), stepped natively over the
    P2P transport: the very kernels of the multi-GPU path (mailboxes, device-side flags, persistent k_u_iter with the
    convergence rounds inside), the peers being plain device pointers.  The bands' kernels wait for each other on the
    device, so they are enqueued tick by tick, round-robin."""

    def __init__(self, bands):
        self.bands = bands
        n = len(bands)
        ptrs = [b.p2p_open(r, n)[1] for r, b in enumerate(bands)]
        for b in bands:
            b.p2p_connect(ptrs=ptrs, bands_on_device=n)

    def step(self, perfect, n_ticks):
        for _ in range(n_ticks):
            for b in self.bands:
                b._c(b.L.flowamok_band_step(b.ctx.h, b.grid.h, int(perfect), 1, None))

    def sync(self):
        for b in self.bands:
            b.ctx.sync()

    @property
    def rounds(self):
        return max(b.step_native(False, 0) for b in self.bands)
