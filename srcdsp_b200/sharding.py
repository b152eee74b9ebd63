"""Work partitioning across the GPUs of one box: no collective on this path.

  * independent channels  -> contiguous channel batches per rank (`channel_shard`)
  * one very long stream  -> time slices with a warm-up halo (`time_slices`): slice g starts at
    a multiple of the total decimation; it is preceded by a warm-up block (a multiple of the
    total decimation that covers the filters' delay lines) whose outputs are discarded, so the
    carried history and the NCO phase at the slice start are exactly what a sequential run of
    the reference would have had (dsptl_dnsampling_filters.h:198-205,218-219; mixers.h:177).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence


def channel_shard(channels: int, world: int, rank: int) -> range:
    """Contiguous, balanced channel batch of `rank` (first `channels % world` ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(channels, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def chain_halo(ntaps: Sequence[int], ratios: Sequence[int]) -> int:
    """Input samples that influence the first output of a slice besides the slice itself:
    (N1-1) + (N2-1)*M1 + (N3-1)*M1*M2 ..."""
    halo, scale = 0, 1
    for nt, m in zip(ntaps, ratios):
        halo += (nt - 1) * scale
        scale *= m
    return halo


@dataclass
class TimeSlice:
    rank: int
    start: int        # first input sample of the slice (multiple of the total decimation)
    length: int       # input samples in the slice (multiple of the total decimation)
    warmup: int       # input samples fed before `start` with outputs discarded (0 for slice 0)
    out_start: int    # first output sample index
    out_length: int

    def nco_phase(self, phi0: int, freq: int, n_table: int) -> int:
        """Phase at the first warm-up sample: (phi0 + (start - warmup) * freq) mod N."""
        return (phi0 + ((self.start - self.warmup) % n_table) * freq) % n_table


def time_slices(n_samples: int, world: int, ntaps: Sequence[int], ratios: Sequence[int]) -> List[TimeSlice]:
    total = 1
    for m in ratios:
        total *= m
    if n_samples % total:
        raise ValueError("stream length must be a multiple of the total decimation")
    halo = chain_halo(ntaps, ratios)
    warm = -(-halo // total) * total
    n_out = n_samples // total
    out = []
    for r in range(world):
        o0 = (n_out * r) // world
        o1 = (n_out * (r + 1)) // world
        start, length = o0 * total, (o1 - o0) * total
        w = min(warm, start)
        out.append(TimeSlice(r, start, length, w, o0, o1 - o0))
    return out


def bind_to_gpu_numa_node(device: int) -> dict:
    """Pins the calling process to the CPUs of the NUMA node the GPU hangs off, so that pinned staging buffers
    allocated afterwards live in that node's memory and host<->device copies do not cross the socket link.  With
    one process per GPU and 8 GPUs on a two-socket host this is what keeps the end-to-end (PCIe) rate from
    collapsing when all ranks stream at once.  Best effort: returns what it did, never raises."""
    import os
    info = {"device": device, "numa_node": None, "cpus": None}
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        info["pci"] = bus
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["numa_node"], info["cpus"] = node, len(cpus)
    except Exception as ex:  # no sysfs, no permission, old torch: carry on unbound
        info["error"] = str(ex)[:120]
    return info
