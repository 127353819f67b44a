"""Multi-GPU partitioning of the chain (one process per GPU, torch.distributed).

The path shards two ways and needs at most one tiny exchange:

* by VFO row (``--simo``): rows are independent given the raw chunk; each rank takes a contiguous
  slice of the rows and the raw bytes of every batch are broadcast from the ingest rank (NCCL over
  NVLink on the GPU box; the bytes, not the decoded complex128 -- 4x less traffic for int16);
* by time segment (offline files): chunks are independent except for the IQ corrector's offset
  (src/misc/read_file.py:53), an affine recurrence ``off' = lam off + L z``.  Each rank runs its
  segment from a zero offset, the ranks all-gather the offset each segment *gained*, and rank r
  starts from ``sum_{i<r} lam^(n_{i+1..r-1}) g_i``.

Pure host logic; the same functions drive bench.py on NCCL and the gloo tests on CPU."""
from __future__ import annotations

import numpy as np


def row_shard(nrows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [start, stop) of the rows owned by ``rank``, balanced: the row counts of two
    ranks differ by at most one (17 rows over 8 ranks: 3,2,2,2,2,2,2,2), the larger shares first."""
    base, extra = divmod(nrows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def segment_shard(nchunks: int, world: int, rank: int) -> tuple[int, int]:
    """(first chunk, chunk count) of rank's time segment; segments are whole chunks, contiguous,
    in rank order, sizes differing by at most one chunk."""
    base, extra = divmod(nchunks, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def iq_start_offset(gains, samples_per_segment, lam: float, rank: int) -> complex:
    """Offset at the start of ``rank``'s segment from the per-segment gains (each measured from a
    zero offset): fold ``off <- lam^n_i off + g_i`` over the segments before it."""
    off = 0j
    for i in range(rank):
        off = (lam ** int(samples_per_segment[i])) * off + complex(gains[i])
    return off


def exchange_iq_gain(dist, torch, gain: complex, nsamples: int, lam: float, device=None) -> complex:
    """All-gather (gain, segment length) over the default process group and return this rank's
    start offset.  2 complex128 + 1 count per rank on the wire."""
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([gain.real, gain.imag, float(nsamples)], dtype=torch.float64, device=device)
    allv = torch.empty(world * 3, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(allv, mine)
    a = allv.cpu().numpy().reshape(world, 3)
    return iq_start_offset(a[:, 0] + 1j * a[:, 1], a[:, 2], lam, rank)


def broadcast_raw(dist, torch, raw, src: int = 0):
    """Broadcast one raw batch (uint8 tensor, same size on every rank) from the ingest rank."""
    dist.broadcast(raw, src=src)
    return raw


def concat_rows(parts: list[np.ndarray]) -> np.ndarray:
    """Host writer's view: row slices in rank order -> (R, n) (no device gather is ever needed,
    each rank writes its own rows to its own sockets)."""
    return np.concatenate([p for p in parts if p.size], axis=0)
