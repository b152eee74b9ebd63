"""Host-side multi-GPU logic on CPU: channel sharding and time slicing with halo,
including a world_size-2 gloo run that checks the sharded result equals the unsharded one."""
import os
import socket
import sys

import numpy as np
import pytest

import oracle as O
from srcdsp_b200.sharding import chain_halo, channel_shard, time_slices

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_channel_shard_partitions():
    for C in (1, 7, 256, 1024, 1000):
        for W in (1, 2, 3, 4, 8):
            got = [c for r in range(W) for c in channel_shard(C, W, r)]
            assert got == list(range(C))
            sizes = [len(channel_shard(C, W, r)) for r in range(W)]
            assert max(sizes) - min(sizes) <= 1
    assert list(channel_shard(1024, 8, 3)) == list(range(384, 512))


def test_time_slices_cover_and_align():
    sl = time_slices(1 << 20, 8, [1023], [4])
    assert sum(s.length for s in sl) == 1 << 20 and sl[0].warmup == 0
    for a, b in zip(sl, sl[1:]):
        assert a.start + a.length == b.start and b.start % 4 == 0 and b.warmup >= 1022 and b.warmup % 4 == 0
    assert chain_halo([63, 63], [8, 4]) == 62 + 62 * 8
    sl = time_slices(32 * 1000, 3, [63, 63], [8, 4])
    assert all(s.start % 32 == 0 and s.length % 32 == 0 and s.warmup % 32 == 0 for s in sl)


def _sliced_ddc(corc, x, slices, f, t1, M1, t2, M2):
    """Each slice is processed independently (as one rank would): warm-up block, then the slice."""
    outs = []
    fr = corc.mixer_set_frequency(f)
    for s in slices:
        phi = s.nco_phase(0, fr, 4096)
        h1 = h2 = None
        lo = s.start - s.warmup
        if s.warmup:
            y, phi = corc.mixer_step(x[lo:s.start], phi, fr)
            y, h1 = corc.dec_step(t1, M1, y, h1)
            _, h2 = corc.dec_step(t2, M2, y, h2)
        y, phi = corc.mixer_step(x[s.start:s.start + s.length], phi, fr)
        y, h1 = corc.dec_step(t1, M1, y, h1)
        y, h2 = corc.dec_step(t2, M2, y, h2)
        assert y.shape[0] == s.out_length
        outs.append(y)
    return np.concatenate(outs)


def test_time_sliced_chain_equals_sequential(corc):
    """Slicing one long stream with a warm-up halo is bit-exact (the contract the GPU time-slice
    path relies on), checked on the oracle."""
    n = 32 * 600
    x = corc.synth(0x5EED0005, 0, 0, n, 0)
    t1, t2 = O.design_lowpass_taps(63, 8), O.design_lowpass_taps(63, 4)
    f = -0.3217
    whole = _sliced_ddc(corc, x, time_slices(n, 1, [63, 63], [8, 4]), f, t1, 8, t2, 4)
    for W in (2, 3, 8):
        assert np.array_equal(_sliced_ddc(corc, x, time_slices(n, W, [63, 63], [8, 4]), f, t1, 8, t2, 4), whole)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle as OO
    from srcdsp_b200.sharding import channel_shard as cs
    dist.init_process_group("gloo", rank=rank, world_size=world)
    C, n, M = 6, 8 * 64, 8
    taps = OO.design_lowpass_taps(63, 8)
    c = OO.corc()
    mine = cs(C, world, rank)
    chk = 0
    for ch in mine:  # each rank filters only its own channels: no data-path collective
        y, _ = c.dec_step(taps, M, c.synth(0x5EED0003, ch, 0, n, 0))
        chk += int(y.astype(np.int64).sum()) + 31 * ch
    # the only communication is the measurement plumbing bench.py uses: barrier + max/sum reduce
    dist.barrier()
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    units = torch.tensor([len(mine) * (n // M)], dtype=torch.int64)
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    s = torch.tensor([chk], dtype=torch.int64)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put((t.item(), units.item(), s.item()))
    dist.destroy_process_group()


def test_world_size_2_gloo_channel_sharding(corc):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    tmax, units, chk = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    C, n, M = 6, 8 * 64, 8
    taps = O.design_lowpass_taps(63, 8)
    ref = 0
    for ch in range(C):
        y, _ = corc.dec_step(taps, M, corc.synth(0x5EED0003, ch, 0, n, 0))
        ref += int(y.astype(np.int64).sum()) + 31 * ch
    assert tmax == 2.0 and units == C * (n // M) and chk == ref


def _slice_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle as OO
    from srcdsp_b200.sharding import time_slices as ts
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, M, nt = 4 * 5000, 4, 1023
    taps = OO.design_lowpass_taps(nt, M)
    c = OO.corc()
    sl = ts(n, world, [nt], [M])[rank]
    # what bench.py --workload cfg5 does per rank: the rank holds [start - warmup, start + length) of the ONE stream
    # (counter-based synthesis: no rank needs another rank's samples), warm-up with outputs discarded, then its slice
    x = c.synth(0x5EED0005, 0, sl.start - sl.warmup, sl.warmup + sl.length, 0)
    h = None
    if sl.warmup:
        _, h = c.dec_step(taps, M, x[:sl.warmup], h)
    y, _ = c.dec_step(taps, M, x[sl.warmup:], h)
    assert y.shape[0] == sl.out_length
    # outputs land at [out_start, out_start + out_length) of the stream's output; summing the disjoint pieces is
    # the test's stand-in for "one pinned host buffer at per-device offsets" (no data-path collective in the product)
    full = torch.zeros((n // M, 2), dtype=torch.int32)
    full[sl.out_start:sl.out_start + sl.out_length] = torch.from_numpy(y.astype(np.int32))
    dist.all_reduce(full, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put(full.numpy().astype(np.int16))
    dist.destroy_process_group()


def test_world_size_2_gloo_time_slices(corc):
    """cfg5's partition on two ranks: slice + warm-up halo per rank == the sequential run of the whole stream."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_slice_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, M, nt = 4 * 5000, 4, 1023
    whole, _ = corc.dec_step(O.design_lowpass_taps(nt, M), M, corc.synth(0x5EED0005, 0, 0, n, 0))
    assert np.array_equal(got, whole)


def test_time_slices_properties_randomised():
    """For random chains, stream lengths and world sizes: the slices tile the stream, start at multiples of the total
    decimation, and each one's warm-up either reaches back over the whole delay-line reach or to the stream's start
    (where the reset filter's zero history IS the sequential run's) -- the two cases in which slicing is exact."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.tuples(st.integers(1, 300), st.integers(1, 16)), min_size=1, max_size=3),
           st.integers(1, 4000), st.integers(1, 9))
    def check(chain, blocks, world):
        ntaps, ratios = [c[0] for c in chain], [c[1] for c in chain]
        total = int(np.prod(ratios))
        n = blocks * total
        sl = time_slices(n, world, ntaps, ratios)
        halo = chain_halo(ntaps, ratios)
        assert len(sl) == world and sl[0].start == 0 and sl[0].warmup == 0
        assert sum(s.length for s in sl) == n and sum(s.out_length for s in sl) == n // total
        pos = 0
        for s in sl:
            assert s.start == pos and s.start % total == 0 and s.length % total == 0 and s.warmup % total == 0
            assert s.out_start * total == s.start and s.out_length * total == s.length
            assert 0 <= s.warmup <= s.start and (s.warmup >= halo or s.warmup == s.start)
            assert s.nco_phase(5, 3437, 4096) == (5 + (s.start - s.warmup) * 3437) % 4096
            pos += s.length
        sizes = [s.out_length for s in sl]
        assert max(sizes) - min(sizes) <= 1  # balanced to one output

    check()
    with pytest.raises(ValueError):
        time_slices(33, 2, [63], [8])
