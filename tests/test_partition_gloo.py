"""CPU: the N>1 host logic -- row-group partitioning + the count/aggregate all-reduce -- with a
world_size-2 gloo job.  Each rank evaluates its own row-group range with the ORACLE (there is
no GPU here and the product has no CPU path); the reduced counts must equal a single-process
evaluation of the whole column."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def test_row_group_ranges():
    import fle_b200  # noqa: F401
    from fle_b200.partition import rank_rows, row_group_ranges
    for nrg in (0, 1, 7, 8, 9, 1431):
        for w in (1, 2, 4, 8):
            r = row_group_ranges(nrg, w)
            assert r[0][0] == 0 and r[-1][1] == nrg
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    # SF1000 lineitem: 5,999,989,709 rows in 4 Mi-row groups over 8 GPUs
    total = 5_999_989_709
    spans = [rank_rows(total, 8, r) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert all(lo % 64 == 0 for lo, _ in spans)


def _worker(rank, world, port, n, rg_rows, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fle_b200  # noqa: F401
    from fle_b200.partition import all_reduce_sums, rank_rows
    from oracle.oracle import AND, GE, LE, LT, Oracle
    orc = Oracle()
    rng = np.random.default_rng(123)  # every rank generates the same table, scans only its rows
    cols = [rng.integers(0, d, size=n).astype(np.uint32) for d in (2526, 11, 50)]
    price = rng.integers(90_000, 10_500_000, size=n).astype(np.int64)
    lo, hi = rank_rows(n, world, rank, rg_rows)
    enc = [(orc.pack(c[lo:hi], bw), bw) for c, bw in zip(cols, (12, 4, 6))]
    q6 = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
          (AND, None, None), (LT, 2, 23), (AND, None, None)]
    bm, cnt = orc.eval_tree(enc, q6, hi - lo)
    bits = np.unpackbits(bm.view(np.uint8), bitorder="little")[: hi - lo].astype(bool)
    revenue = int((price[lo:hi][bits] * cols[1][lo:hi][bits].astype(np.int64)).sum())  # cents * discount code
    tot_cnt, tot_rev = all_reduce_sums([cnt, revenue])
    q.put((rank, lo, hi, cnt, tot_cnt, tot_rev))
    dist.destroy_process_group()


def test_two_rank_q6_count_and_revenue_reduce():
    n, rg_rows, world = 64 * 700 + 13, 64 * 100, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, rg_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(123)
    cols = [rng.integers(0, d, size=n).astype(np.uint32) for d in (2526, 11, 50)]
    price = rng.integers(90_000, 10_500_000, size=n).astype(np.int64)
    m = (cols[0] >= 730) & (cols[0] <= 1094) & (cols[1] >= 5) & (cols[1] <= 7) & (cols[2] < 23)
    want_cnt, want_rev = int(m.sum()), int((price[m] * cols[1][m].astype(np.int64)).sum())
    res.sort()
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n  # contiguous, complete
    assert sum(r[3] for r in res) == want_cnt
    for r in res:
        assert r[4] == want_cnt and r[5] == want_rev
