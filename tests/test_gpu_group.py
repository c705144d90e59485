"""Row groups partitioned across the members of a group, totals summed INSIDE the scan kernel
(fle_group_*, fle_eval_group_dev, fle_eval_sum_product_group_dev; SURVEY 8e, BASELINE cfg4).

On the one-GPU test box the members are several contexts (one stream each) on cuda:0: the exchange
code is the same peer-store / acquire-poll code that crosses NVLink between GPUs, and the kernels of
different members really run concurrently.  bench.py --gpus N runs the same entry points with one
process per GPU over CUDA IPC mailboxes."""
import numpy as np
import pytest

from oracle.oracle import AND, EQ, GE, IN, LE, LT, OR

pytestmark = pytest.mark.gpu

Q6 = [(GE, 0, 730), (LE, 0, 1094), (AND, None, None), (GE, 1, 5), (LE, 1, 7), (AND, None, None),
      (AND, None, None), (LT, 2, 23), (AND, None, None)]
COLS = ((2526, 12), (11, 4), (50, 6))


def _members(world):
    import torch
    import fle_b200
    from fle_b200 import capi
    streams = [torch.cuda.Stream() for _ in range(world)]
    ctxs = [fle_b200.FleContext(0, s.cuda_stream) for s in streams]
    capi.group_create_local(ctxs)
    return streams, ctxs


def _close(ctxs):
    for c in ctxs:
        st = c.group_status()
        assert st["peers_timed_out"] == 0, st
    for c in ctxs:
        c.group_close()
        c.close()


@pytest.mark.parametrize("deferred", [0, 1])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_cfg4_q6_partitioned_reduced_scale(oracle, world, deferred):
    """cfg4 at reduced scale: rows cut into row groups, contiguous row-group ranges per member
    (partition.rank_rows), fused Q6 + exact int64 revenue, totals identical on every member and equal to
    the oracle's evaluation of the whole table -- for any member count."""
    import torch
    from fle_b200 import partition
    rng = np.random.default_rng(0xF1E00400)
    total_rows, rg_rows = 64 * 9000 + 37, 64 * 500
    vals = [rng.integers(0, d, size=total_rows).astype(np.uint32) for d, _ in COLS]
    price = rng.integers(90_000, 10_500_000, size=total_rows).astype(np.int64)
    whole = [(oracle.pack(v, bw), bw) for v, (_, bw) in zip(vals, COLS)]
    want_bm, want_cnt = oracle.eval_tree(whole, Q6, total_rows)
    sel = np.unpackbits(want_bm.view(np.uint8), bitorder="little")[:total_rows].astype(bool)
    want_rev = int((price[sel] * vals[1][sel].astype(np.int64)).sum())

    streams, ctxs = _members(world)
    for c in ctxs:
        c.set_option("group_deferred", deferred)
    ddict = torch.arange(0, 11, dtype=torch.int64, device="cuda")
    shards, outs, bms = [], [], []
    for r in range(world):
        first, end = partition.rank_rows(total_rows, world, r, rg_rows)
        assert first % 64 == 0
        n = end - first
        dcols = [(torch.from_numpy(oracle.pack(v[first:end], bw).view(np.int64)).cuda(), bw) for v, (_, bw) in zip(vals, COLS)]
        shards.append((first, n, dcols, torch.from_numpy(price[first:end]).cuda()))
    torch.cuda.synchronize()
    # (members share ONE device here: load every kernel before any member waits for its peers -- with
    #  lazy module loading the first launch of a function can wait for the kernels that are running)
    for jit in (2, 0):
        ctxs[0].set_option("eval_jit", jit)
        ctxs[0].eval_sum_product_dev(shards[0][2], Q6, shards[0][1], shards[0][3], 1, ddict)
    torch.cuda.synchronize()
    all_outs = []
    for it in range(12):  # queries back to back, more than there are mailbox slots; deferred: no sync in between
        outs, bms = [], []
        for r, (first, n, dcols, dprice) in enumerate(shards):
            ctxs[r].set_option("eval_jit", 2 if it % 3 < 2 else 0)  # fused single kernel / the two-kernel fallback + reduce kernel
            with torch.cuda.stream(streams[r]):
                bm = torch.zeros(max(1, (n + 63) // 64), dtype=torch.int64, device="cuda")
                out = ctxs[r].eval_sum_product_group_dev(dcols, Q6, n, dprice, 1, ddict, bitmap=bm)
            outs.append(out)
            bms.append(bm)
        all_outs.append(outs)
        if not deferred or it == 11:
            if deferred:
                for r in range(world):
                    with torch.cuda.stream(streams[r]):
                        ctxs[r].group_flush()
            torch.cuda.synchronize()
            for r, (first, n, _, _) in enumerate(shards):
                got = bms[r].cpu().numpy().view(np.uint64)[: (n + 63) // 64]
                assert (got == want_bm[first // 64: first // 64 + len(got)]).all(), "local bitmap differs from the oracle's slice"
    for it, outs in enumerate(all_outs):
        for r in range(world):
            assert int(outs[r][0].item()) == want_rev and int(outs[r][1].item()) == want_cnt, (world, it, r)
    _close(ctxs)


def test_group_counts_all_kernel_families(oracle):
    """The exchange rides on every kernel family's epilogue: single-pass, narrow, IN lookup, specialised
    multi-column, interpreting, and the regrouped long-IN path (trailing reduce kernel)."""
    import torch
    world = 3
    rng = np.random.default_rng(7)
    n_per = [64 * 700 + 5, 64 * 300, 64 * 1200 + 63]
    progs = [
        ([12], [(GE, 0, 100), (LE, 0, 3000), (AND, None, None)]),
        ([2], [(EQ, 0, 1)]),
        ([1], [(EQ, 0, 1)]),
        ([12], [(IN, 0, list(range(5, 900, 7)))]),
        ([12, 6], [(LT, 0, 2000), (EQ, 1, 17), (OR, None, None)]),
        ([12, 6], [(IN, 0, list(range(0, 4000, 3))), (LT, 1, 23), (AND, None, None)]),
    ]
    streams, ctxs = _members(world)
    for jit in (0, 2):
        for bws, prog in progs:
            want_total, runs = 0, []
            for r in range(world):
                vals = [rng.integers(0, 1 << bw, size=n_per[r]).astype(np.uint32) for bw in bws]
                cols = [(oracle.pack(v, bw), bw) for v, bw in zip(vals, bws)]
                wbm, wcnt = oracle.eval_tree(cols, prog, n_per[r])
                want_total += wcnt
                runs.append((wbm, [(torch.from_numpy(w.view(np.int64)).cuda(), bw) for w, bw in cols]))
            ctxs[0].set_option("eval_jit", jit)
            ctxs[0].set_option("in_split_min_rows", 0)
            ctxs[0].eval_dev(runs[0][1], prog, n_per[0])  # loads the kernels (see above)
            torch.cuda.synchronize()
            res = []
            for r in range(world):
                ctxs[r].set_option("eval_jit", jit)
                ctxs[r].set_option("in_split_min_rows", 0)
                with torch.cuda.stream(streams[r]):
                    res.append(ctxs[r].eval_group_dev(runs[r][1], prog, n_per[r]))
            torch.cuda.synchronize()
            for r in range(world):
                bm, cnt = res[r]
                assert int(cnt.item()) == want_total, (bws, prog[0], jit, r)
                assert (bm.cpu().numpy().view(np.uint64)[: len(runs[r][0])] == runs[r][0]).all()
    _close(ctxs)


def test_group_of_one_and_empty_member(oracle):
    import torch
    import fle_b200
    from fle_b200 import capi
    c = fle_b200.FleContext(0, torch.cuda.current_stream().cuda_stream)
    capi.group_create_local([c])
    v = np.arange(64 * 10, dtype=np.uint32) % 50
    cols = [(torch.from_numpy(oracle.pack(v, 6).view(np.int64)).cuda(), 6)]
    _, cnt = c.eval_group_dev(cols, [(LT, 0, 23)], len(v))
    torch.cuda.synchronize()
    assert int(cnt.item()) == int((v < 23).sum())
    c.group_close()
    c.close()
    # a member with zero rows still takes part in the sum
    streams, ctxs = _members(2)
    ctxs[0].eval_dev(cols, [(LT, 0, 23)], len(v))
    torch.cuda.synchronize()
    with torch.cuda.stream(streams[0]):
        _, c0 = ctxs[0].eval_group_dev(cols, [(LT, 0, 23)], len(v))
    with torch.cuda.stream(streams[1]):
        _, c1 = ctxs[1].eval_group_dev(cols, [(LT, 0, 23)], 0)
    torch.cuda.synchronize()
    assert int(c0.item()) == int(c1.item()) == int((v < 23).sum())
    _close(ctxs)
