"""Row-group partitioning across the GPUs of one box, and the one collective the path has.

Row groups are independent (the reference notes they "could be parallelized",
hdfs-parquet-scanner.cc:1056-1060); a predicate never needs a neighbouring row, so every rank
scans a contiguous range of row groups and keeps its bitmap local.  The only exchange is the
SUM of a few 64-bit scalars per query (match count, exact int64 partial aggregates such as the
TPC-H Q6 revenue in cents): one all-reduce over NCCL/NVLink, latency-bound.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

ROW_GROUP_ROWS = 4_194_304  # SF1000-shaped lineitem row groups (multiple of 64)


def row_group_ranges(num_row_groups: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [lo, hi) row-group ranges, one per rank (sizes differ by at most 1)."""
    if world_size < 1 or num_row_groups < 0:
        raise ValueError("bad partition request")
    out = []
    for r in range(world_size):
        lo = num_row_groups * r // world_size
        hi = num_row_groups * (r + 1) // world_size
        out.append((lo, hi))
    return out


def rank_rows(total_rows: int, world_size: int, rank: int, row_group_rows: int = ROW_GROUP_ROWS) -> Tuple[int, int]:
    """[first_row, end_row) of `rank` when total_rows are cut into row groups of row_group_rows
    (the last one ragged) and the row groups are dealt out contiguously."""
    nrg = (total_rows + row_group_rows - 1) // row_group_rows
    lo, hi = row_group_ranges(nrg, world_size)[rank]
    return lo * row_group_rows, min(hi * row_group_rows, total_rows)


def all_reduce_sums(values: Sequence[int], device=None, group=None) -> List[int]:
    """SUM-all-reduce of a few int64 scalars (match counts / partial aggregates).  Uses the
    initialised torch.distributed backend (NCCL on GPUs, gloo in CPU tests); identity when the
    job is a single process."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [int(v) for v in values]
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [int(v) for v in t.tolist()]
