"""GPU parity of the kernels specialised per program shape at run time (fle_jit.cu /
fle_eval_jit.cuh): same inputs through eval_jit = 2 and through the CPU oracle, bit-exact; plus
the policy (auto mode compiles only for large calls, shapes are shared across literals)."""
import numpy as np
import pytest

from oracle.oracle import ALL, AND, EQ, GE, GT, IN, LE, LT, NONE, OR

pytestmark = pytest.mark.gpu

N_ = (None, None)


def q6(lo=730, hi=1094, dlo=5, dhi=7, q=23):
    return [(GE, 0, lo), (LE, 0, hi), (AND,) + N_, (GE, 1, dlo), (LE, 1, dhi), (AND,) + N_, (AND,) + N_, (LT, 2, q),
            (AND,) + N_]


@pytest.fixture()
def jit_ctx(ctx):
    ctx.set_option("eval_variant", 1)
    ctx.set_option("eval_jit", 2)
    yield ctx
    ctx.set_option("eval_jit", 1)
    ctx.set_option("eval_gpt", 0)
    ctx.set_option("eval_stages", 0)


def make_cols(oracle, rng, n, bws, doms=None):
    doms = doms or [1 << b for b in bws]
    vals = [rng.integers(0, d, size=n, dtype=np.uint64).astype(np.uint32) for d in doms]
    return vals, [(oracle.pack(v, bw), bw) for v, bw in zip(vals, bws)]


@pytest.mark.parametrize("n", [1, 63, 64, 65, 64 * 255 + 1, 64 * 256, 64 * 256 * 3 + 5, 64 * 2048 * 2 + 64 * 100 + 7])
def test_ragged_sizes(jit_ctx, oracle, n):
    rng = np.random.default_rng(n)
    _, cols = make_cols(oracle, rng, n, (12, 4, 6), (2526, 11, 50))
    before = jit_ctx.get_option("jit_launches")
    for gpt in (0, 1, 2):
        jit_ctx.set_option("eval_gpt", gpt)
        bm, cnt = jit_ctx.eval(cols, q6(), n)
        want, wcnt = oracle.eval_tree(cols, q6(), n)
        assert (bm == want).all() and cnt == wcnt, (n, gpt)
    assert jit_ctx.get_option("jit_launches") >= before + 3


def test_one_kernel_per_shape_any_literals(jit_ctx, oracle):
    rng = np.random.default_rng(7)
    n = 64 * 5000 + 3
    _, cols = make_cols(oracle, rng, n, (12, 4, 6), (2526, 11, 50))
    progs = [q6(), q6(0, 2525, 0, 10, 49), q6(1000, 10, 7, 3, 0),
             [(LT, 0, 100), (GT, 0, 2000), (AND,) + N_, (LE, 1, 3), (GE, 1, 1), (AND,) + N_, (AND,) + N_, (GE, 2, 30),
              (AND,) + N_]]
    for prog in progs:
        bm, cnt = jit_ctx.eval(cols, prog, n)
        want, wcnt = oracle.eval_tree(cols, prog, n)
        assert (bm == want).all() and cnt == wcnt


def test_trees_consts_in_lists(jit_ctx, oracle):
    rng = np.random.default_rng(11)
    n = 64 * 9000 + 41
    vals, cols = make_cols(oracle, rng, n, (12, 4, 6, 3, 16, 1, 31, 17), (2526, 11, 50, 7, 40000, 2, 1 << 31, 1 << 17))
    long_in = rng.choice(vals[0], size=20).astype(np.uint64)
    progs = {
        "or_tree": (cols[:3], [(IN, 1, [2, 9]), (EQ, 2, 17), (OR,) + N_, (LT, 0, 1200), (AND,) + N_, (GT, 1, 8), (OR,) + N_]),
        "flat_or_long_in_first": ([cols[0], cols[2]], [(IN, 0, long_in), (IN, 0, [5, 6, 7]), (OR,) + N_, (LT, 1, 3), (OR,) + N_]),
        "flat_and_long_in": ([cols[0], cols[2]], [(IN, 0, long_in), (LT, 1, 30), (AND,) + N_]),
        "empty_in": ([cols[0], cols[2]], [(IN, 0, np.zeros(0, np.uint64)), (LT, 1, 30), (OR,) + N_]),
        "consts": (cols[:2], [(ALL,) + N_, (LT, 0, 1000), (AND,) + N_, (NONE,) + N_, (OR,) + N_, (GE, 1, 3), (AND,) + N_]),
        "deep": (cols[:3], [(LT, 0, 500), (GT, 0, 2000), (OR,) + N_, (EQ, 1, 3), (EQ, 1, 4), (OR,) + N_, (AND,) + N_,
                            (GE, 2, 10), (LE, 2, 40), (AND,) + N_, (OR,) + N_, (LT, 1, 9), (AND,) + N_]),
        "odd_widths": (cols[5:8], [(EQ, 0, 1), (GT, 1, 1 << 30), (AND,) + N_, (LE, 2, 70000), (AND,) + N_]),
        "eight_columns": (cols, [(LT, 0, 2000)] + sum([[(op, c, lit), (AND,) + N_] for c, (op, lit) in enumerate(
            [(GE, 1), (LE, 45), (GE, 1), (LT, 39000), (EQ, 1), (LT, 1 << 30), (GT, 9)], start=1)], [])),
        "same_column_many_ops": ([cols[0], cols[1]], [(GE, 0, 10), (LE, 0, 2500), (AND,) + N_, (GT, 0, 11), (AND,) + N_,
                                                       (LT, 0, 2499), (AND,) + N_, (EQ, 0, 77), (OR,) + N_, (LT, 1, 5),
                                                       (AND,) + N_]),
    }
    for name, (pc, prog) in progs.items():
        before = jit_ctx.get_option("jit_launches")
        bm, cnt = jit_ctx.eval(pc, prog, n)
        want, wcnt = oracle.eval_tree(pc, prog, n)
        assert (bm == want).all() and cnt == wcnt, name
        assert jit_ctx.get_option("jit_launches") > before, name
        # rows [128, n): the oracle evaluates the same rows on columns cut at the group boundary
        bm2, cnt2 = jit_ctx.eval(pc, prog, n - 128, first_row=128)
        cut = [(w[2 * bw:], bw) for w, bw in pc]
        want2, wcnt2 = oracle.eval_tree(cut, prog, n - 128)
        assert (bm2 == want2).all() and cnt2 == wcnt2, name


def test_auto_mode_policy(ctx, oracle):
    """eval_jit = 1: a shape keeps to the interpreting kernels until the rows evaluated with it
    reach eval_jit_min_rows; then it is compiled and every later call uses the kernel."""
    rng = np.random.default_rng(13)
    n = 64 * 3000
    _, cols = make_cols(oracle, rng, n, (9, 7), (500, 100))  # a shape no other test uses
    prog = [(LT, 0, 250), (GE, 1, 50), (AND,) + N_]
    want, wcnt = oracle.eval_tree(cols, prog, n)
    ctx.set_option("eval_variant", 1)
    ctx.set_option("eval_jit", 1)
    try:
        j0 = ctx.get_option("jit_launches")
        bm, cnt = ctx.eval(cols, prog, n)
        assert (bm == want).all() and cnt == wcnt and ctx.get_option("jit_launches") == j0
        ctx.set_option("eval_jit_min_rows", 3 * n)  # cumulative: this is the shape's third call
        bm, cnt = ctx.eval(cols, prog, n)
        assert (bm == want).all() and cnt == wcnt and ctx.get_option("jit_launches") == j0
        bm, cnt = ctx.eval(cols, prog, n)
        assert (bm == want).all() and cnt == wcnt and ctx.get_option("jit_launches") > j0
        ctx.set_option("eval_jit_min_rows", 1 << 24)
        j1 = ctx.get_option("jit_launches")
        bm, cnt = ctx.eval(cols, prog, n)  # already compiled -> used regardless of size
        assert (bm == want).all() and cnt == wcnt and ctx.get_option("jit_launches") > j1
        ctx.set_option("eval_jit", 0)
        j2 = ctx.get_option("jit_launches")
        bm, cnt = ctx.eval(cols, prog, n)
        assert (bm == want).all() and cnt == wcnt and ctx.get_option("jit_launches") == j2
    finally:
        ctx.set_option("eval_jit", 1)
        ctx.set_option("eval_jit_min_rows", 1 << 24)


def test_background_compilation(ctx, oracle):
    """eval_jit = 3: the shape is compiled on a thread of the library's own; the calling thread never waits for NVRTC
    -- it keeps to the interpreting kernels (same bits) until the kernel is there, then uses it."""
    import time
    rng = np.random.default_rng(14)
    n = 64 * 3000
    _, cols = make_cols(oracle, rng, n, (10, 5, 3), (700, 20, 6))  # a shape no other test uses
    prog = [(LT, 0, 350), (GE, 1, 5), (OR,) + N_, (EQ, 2, 4), (AND,) + N_]
    want, wcnt = oracle.eval_tree(cols, prog, n)
    ctx.set_option("eval_variant", 1)
    ctx.set_option("eval_jit", 3)
    ctx.set_option("eval_jit_min_rows", 1)
    try:
        j0 = ctx.get_option("jit_launches")
        t0 = time.time()
        bm, cnt = ctx.eval(cols, prog, n)  # starts the compilation, runs interpreted
        first_call_s = time.time() - t0
        assert (bm == want).all() and cnt == wcnt and ctx.get_option("jit_launches") == j0
        deadline = time.time() + 60
        while ctx.get_option("jit_launches") == j0 and time.time() < deadline:
            bm, cnt = ctx.eval(cols, prog, n)
            assert (bm == want).all() and cnt == wcnt
            time.sleep(0.02)
        assert ctx.get_option("jit_launches") > j0, "the background compilation never delivered a kernel"
        assert first_call_s < 0.25, f"the first call waited {first_call_s:.2f} s (for the compiler?)"
    finally:
        ctx.set_option("eval_jit", 1)
        ctx.set_option("eval_jit_min_rows", 1 << 24)


def test_device_resident_q6_matches_interpreting_kernel(ctx):
    """SF-10-sized Q6 on device-resident columns: auto mode picks the specialised kernel; its bitmap
    equals the interpreting kernel's word for word (which the other tests pin to the oracle)."""
    import torch
    n = 60_000_013
    g = torch.Generator(device="cuda").manual_seed(3)
    cols = []
    for d, bw in ((2526, 12), (11, 4), (50, 6)):
        v = torch.randint(0, d, (n,), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
        cols.append((ctx.pack_dev(v, bw), bw))
        del v
    G = (n + 63) // 64
    a = torch.zeros(G, dtype=torch.int64, device="cuda")
    b = torch.ones(G, dtype=torch.int64, device="cuda")
    ca = torch.zeros(1, dtype=torch.int64, device="cuda")
    cb = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.set_option("eval_variant", 1)
    ctx.set_option("eval_jit", 1)
    j0 = ctx.get_option("jit_launches")
    ctx.eval_dev(cols, q6(), n, bitmap=a, count=ca)
    assert ctx.get_option("jit_launches") == j0 + 1
    ctx.set_option("eval_jit", 0)
    ctx.eval_dev(cols, q6(), n, bitmap=b, count=cb)
    ctx.set_option("eval_jit", 1)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and int(ca.item()) == int(cb.item()) > 0


def random_program(rng, bws, max_leaves=8):
    """A random postfix tree over the columns: compare / IN / constant leaves, AND / OR nodes."""
    nleaves = int(rng.integers(1, max_leaves + 1))
    prog, depth = [], 0
    for i in range(nleaves):
        c = int(rng.integers(0, len(bws)))
        dom = 1 << bws[c]
        kind = rng.choice(["cmp", "cmp", "cmp", "in", "const"], p=[0.3, 0.3, 0.2, 0.15, 0.05])
        if kind == "cmp":
            prog.append((int(rng.choice([EQ, LT, LE, GT, GE])), c, int(rng.integers(0, dom))))
        elif kind == "in":
            k = int(rng.choice([0, 1, 2, 3, 8, 9, 12]))
            prog.append((IN, c, rng.integers(0, dom, size=k).astype(np.uint64)))
        else:
            prog.append((int(rng.choice([ALL, NONE])),) + N_)
        depth += 1
        while depth > 1 and (i == nleaves - 1 or rng.random() < 0.5):
            prog.append((int(rng.choice([AND, OR])),) + N_)
            depth -= 1
    return prog


@pytest.mark.parametrize("seed", range(12))
def test_random_programs_all_paths_vs_oracle(ctx, oracle, seed):
    """Fuzz: the specialised kernel, the staged interpreter and the direct interpreter against the
    oracle on the same random program (random shapes, widths 1..32, ragged row counts)."""
    rng = np.random.default_rng(1000 + seed)
    ncols = int(rng.integers(1, 5))
    bws = [int(rng.choice([1, 2, 3, 5, 8, 12, 13, 16, 17, 24, 32])) for _ in range(ncols)]
    n = int(rng.choice([64 * 700 + 13, 64 * 4096 + 1, 64 * 5000]))
    _, cols = make_cols(oracle, rng, n, bws, [min(1 << b, 50) if rng.random() < 0.5 else 1 << b for b in bws])
    try:
        for _ in range(3):
            prog = random_program(rng, bws)
            want, wcnt = oracle.eval_tree(cols, prog, n)
            for variant, jit in ((1, 2), (2, 0), (0, 0)):
                ctx.set_option("eval_variant", variant)
                ctx.set_option("eval_jit", jit)
                bm, cnt = ctx.eval(cols, prog, n)
                assert (bm == want).all() and cnt == wcnt, (seed, variant, jit, prog)
    finally:
        ctx.set_option("eval_variant", 1)
        ctx.set_option("eval_jit", 1)


@pytest.mark.parametrize("jit", [1, 0])
@pytest.mark.parametrize("n", [64 * 50000 + 41, 64 * 256 * 2, 37])
def test_fused_scan_and_aggregate(ctx, jit, n):
    """fle_eval_sum_product_dev (predicate + SUM(price * dict[discount]) in one specialised kernel)
    against torch on the decoded columns -- exact int64 -- and against the two-kernel path (jit=0)."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(n)
    ship = torch.randint(0, 2526, (n,), generator=g, device="cuda", dtype=torch.int32)
    disc = torch.randint(0, 11, (n,), generator=g, device="cuda", dtype=torch.int32)
    qty = torch.randint(0, 50, (n,), generator=g, device="cuda", dtype=torch.int32)
    price = torch.randint(90_000, 10_500_000, (n,), generator=g, device="cuda", dtype=torch.int64)
    cols = [(ctx.pack_dev(ship, 12), 12), (ctx.pack_dev(disc, 4), 4), (ctx.pack_dev(qty, 6), 6)]
    dict_disc = torch.arange(0, 11, device="cuda", dtype=torch.int64) * 100 + 7
    ctx.set_option("eval_jit", jit)
    try:
        for lo, hi in ((730, 1094), (0, 2525), (2000, 10)):
            prog = q6(lo, hi)
            mask = (ship >= lo) & (ship <= hi) & (disc >= 5) & (disc <= 7) & (qty < 23)
            want = int((price[mask] * dict_disc[disc[mask].long()]).sum().item())
            j0 = ctx.get_option("jit_launches")
            bm = torch.zeros((n + 63) // 64, dtype=torch.int64, device="cuda")
            out = ctx.eval_sum_product_dev(cols, prog, n, price, 1, dict_disc, bitmap=bm)
            torch.cuda.synchronize()
            assert int(out[0].item()) == want and int(out[1].item()) == int(mask.sum().item()), (lo, hi)
            assert (ctx.get_option("jit_launches") > j0) == (jit != 0)
            bm2, cnt2 = ctx.eval_dev(cols, prog, n)
            assert torch.equal(bm, bm2)
            out2 = ctx.eval_sum_product_dev(cols, prog, n, price, 1, dict_disc)  # no bitmap wanted
            torch.cuda.synchronize()
            assert torch.equal(out, out2)
        # rows [128, n) of the same columns; a dictionary shorter than the codes (code >= ndict counts 0)
        if n > 128:
            mask = (ship >= 730) & (ship <= 1094) & (disc >= 5) & (disc <= 7) & (qty < 23)
            mask[:128] = False
            short = dict_disc[:7]
            dd = torch.where(disc < 7, short[disc.clamp(max=6).long()], torch.zeros_like(price))
            want = int((price[mask] * dd[mask]).sum().item())
            out = ctx.eval_sum_product_dev(cols, q6(), n - 128, price[128:], 1, short, first_row=128)
            torch.cuda.synchronize()
            assert int(out[0].item()) == want and int(out[1].item()) == int(mask.sum().item())
    finally:
        ctx.set_option("eval_jit", 1)
