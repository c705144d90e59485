"""Row-group batch kernels behind the scanner's AssembleRows (fle_pages_gather_dev, the single-pass
fle_bitmap_select_dev, fle_assemble_rows_dev) against numpy restatements of the reference loops
(hdfs-parquet-scanner.cc:1134-1180, dict-encoding.h:185-192, :321-330, :438-444)."""
import numpy as np
import pytest

from oracle.oracle import GE

pytestmark = pytest.mark.gpu


def _bits(words, n):
    return np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")[:n].astype(bool)


@pytest.mark.parametrize("n", [1, 63, 64, 65536, 65537, 64 * 70000 + 17, 1 << 24])
@pytest.mark.parametrize("density", [0.0, 0.018, 0.5, 1.0])
def test_single_pass_select(ctx, n, density):
    import torch
    rng = np.random.default_rng(n % 1000 + int(density * 100))
    sel = rng.random(n) < density if 0 < density < 1 else np.full(n, density == 1.0)
    words = np.packbits(np.concatenate([sel, np.zeros((-n) % 64, dtype=bool)]), bitorder="little").view(np.uint64)
    words = words.copy()
    if n % 64:
        words[-1] |= np.uint64(0xFFFFFFFFFFFFFFFF) << np.uint64(n % 64)  # garbage past the end must be ignored
    bm = torch.from_numpy(words.view(np.int64)).cuda()
    want = np.nonzero(sel)[0].astype(np.uint32)
    for rep in range(3):  # launches back to back: the tile status words are reused, only the generation changes
        ids, cnt = ctx.bitmap_select_dev(bm, n)
    torch.cuda.synchronize()
    k = int(cnt.item())
    assert k == len(want)
    assert (ids.cpu().numpy().view(np.uint32)[:k] == want).all()
    # capacity smaller than the number of survivors: the count is still exact, the prefix is written
    cap = max(1, k // 3)
    ids2, cnt2 = ctx.bitmap_select_dev(bm, n, cap=cap)
    torch.cuda.synchronize()
    assert int(cnt2.item()) == k
    assert (ids2.cpu().numpy().view(np.uint32)[:min(cap, k)] == want[:min(cap, k)]).all()


def test_single_pass_select_long_bitmap(ctx):
    """Bitmaps long enough for the 65536-row tiles (shorter ones take 16384-row tiles so that every SM gets work)."""
    import torch
    n = (1 << 25) + 3
    g = torch.Generator(device="cuda").manual_seed(9)
    sel = torch.rand(n, generator=g, device="cuda") < 0.018
    G = (n + 63) // 64
    pad = torch.cat([sel, torch.zeros(G * 64 - n, dtype=torch.bool, device="cuda")]).view(G, 64)
    bm = torch.zeros(G, dtype=torch.int64, device="cuda")
    for b in range(64):
        bm |= pad[:, b].to(torch.int64) << b
    want = torch.nonzero(sel).flatten().to(torch.int32)
    ids, cnt = ctx.bitmap_select_dev(bm, n)
    torch.cuda.synchronize()
    k = int(cnt.item())
    assert k == want.numel()
    assert torch.equal(ids[:k].to(torch.int32), want)


def test_pages_gather_widths_offsets_and_nulls(ctx, oracle):
    """Pages of one column chunk at odd offsets of the chunk image, written with growing bit widths
    (the reference's writer picks the width per page), gathered into one plane array at the widest width."""
    import torch
    rng = np.random.default_rng(11)
    page_rows = [64 * 300, 64 * 300, 64 * 300, 64 * 120 + 37]
    page_bw = [3, 5, 5, 7]
    dst_bw = 7
    blob = bytearray()
    refs, all_vals, nulls_want = [], [], 0
    for rows, bw in zip(page_rows, page_bw):
        vals = rng.integers(0, 1 << bw, size=rows).astype(np.uint32)
        all_vals.append(vals)
        words = oracle.pack(vals, bw)
        defs = rng.random(rows) < 0.97
        nulls_want += int((~defs).sum())
        dwords = oracle.pack(defs.astype(np.uint32), 1)
        blob += bytes(rng.integers(0, 255, size=int(rng.integers(17, 40)), dtype=np.uint8))  # "page header"
        blob += np.int32(dwords.nbytes).tobytes()
        def_off = len(blob)
        blob += dwords.tobytes()
        blob += bytes([bw])
        data_off = len(blob)
        blob += words.tobytes()
        refs.append(dict(src_off=data_off, src_bytes=words.nbytes, src_bit_width=bw, groups=(rows + 63) // 64,
                         def_off=def_off, def_bytes=dwords.nbytes, rows=rows))
    blob += bytes(16)
    dblob = torch.from_numpy(np.frombuffer(bytes(blob), dtype=np.uint8).copy()).cuda()
    total_groups = sum(r["groups"] for r in refs)
    planes = torch.full((total_groups * dst_bw,), -1, dtype=torch.int64, device="cuda")
    g0, crefs = 0, []
    for r in refs:
        crefs.append(dict(src_dev=dblob.data_ptr() + r["src_off"], src_bytes=r["src_bytes"], src_bit_width=r["src_bit_width"],
                          dst_bit_width=dst_bw, dst_dev=planes.data_ptr() + g0 * dst_bw * 8, groups=r["groups"],
                          def_src_dev=dblob.data_ptr() + r["def_off"], def_bytes=r["def_bytes"], rows=r["rows"]))
        g0 += r["groups"]
    nulls = ctx.pages_gather_dev(crefs)
    torch.cuda.synchronize()
    assert int(nulls.item()) == nulls_want
    got = planes.cpu().numpy().view(np.uint64)
    g0 = 0
    for r, vals in zip(refs, all_vals):
        dec = oracle.unpack(got[g0 * dst_bw:(g0 + r["groups"]) * dst_bw], dst_bw, 0, len(vals))
        assert (dec == vals).all()
        g0 += r["groups"]
    # ... and the fused kernel scans the gathered column as one plane array (ragged last page: rows are per page)
    n_full = sum(page_rows[:3])
    bm, cnt = ctx.eval_dev([(planes, dst_bw)], [(GE, 0, 3)], n_full)
    torch.cuda.synchronize()
    assert int(cnt.item()) == int((np.concatenate(all_vals[:3]) >= 3).sum())


@pytest.mark.parametrize("slot_sizes", [(4, 4, 4), (8, 4, 2, 1), (1,), (8, 8)])
def test_assemble_rows(ctx, oracle, slot_sizes):
    import torch
    rng = np.random.default_rng(len(slot_sizes))
    n = 64 * 5000 + 21
    dtypes = {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}
    sel = rng.random(n) < 0.03
    words = np.packbits(np.concatenate([sel, np.zeros((-n) % 64, dtype=bool)]), bitorder="little").view(np.uint64)
    bm = torch.from_numpy(words.view(np.int64).copy()).cuda()
    ncols = len(slot_sizes)
    off = ncols
    slots, keep, cols = [], [], []
    for c, ss in enumerate(slot_sizes):
        bw = int(rng.integers(1, 14))
        nd = (1 << bw) - (1 if c == 0 and bw > 2 else 0)  # one column with a code outside its dictionary
        codes = rng.integers(0, 1 << bw, size=n).astype(np.uint32)
        d = rng.integers(-100, 100, size=nd).astype(dtypes[ss])
        off = (off + ss - 1) // ss * ss
        planes = torch.from_numpy(oracle.pack(codes, bw).view(np.int64)).cuda()
        dd = torch.from_numpy(d).cuda()
        keep += [planes, dd]
        slots.append(dict(planes_dev=planes.data_ptr(), bytes=planes.numel() * 8, bit_width=bw, slot_size=ss,
                          dict_dev=dd.data_ptr(), ndict=nd, tuple_offset=off))
        cols.append((codes, d, off, ss, nd))
        off += ss
    tuple_size = (off + 7) // 8 * 8
    k = int(sel.sum())
    tuples, cnt, bad = ctx.assemble_rows_dev(bm, n, slots, tuple_size, k + 5)
    torch.cuda.synchronize()
    assert int(cnt.item()) == k
    got = tuples.cpu().numpy()[:k]
    want = np.zeros((k, tuple_size), dtype=np.uint8)
    rows = np.nonzero(sel)[0]
    bad_want = 0
    for codes, d, o, ss, nd in cols:
        cs = codes[rows]
        ok = cs < nd
        bad_want += int((~ok).sum())
        v = np.zeros(k, dtype=d.dtype)
        v[ok] = d[cs[ok]]
        want[:, o:o + ss] = v.view(np.uint8).reshape(k, ss)
    assert int(bad.item()) == bad_want
    assert (got == want).all()


@pytest.mark.parametrize("null_frac", [0.0, 0.03, 0.5, 1.0])
def test_pages_gather_expands_null_pages(ctx, oracle, null_frac):
    """Pages whose FLE section holds only the non-null values (hdfs-parquet-scanner.cc:886-921) are expanded into
    row-aligned planes: the value at every defined row, code 0 at NULL rows; the definition levels travel as a 1-bit
    column of their own.  Ragged last page, per-page widths, odd offsets, an all-NULL page without a data section."""
    import torch
    rng = np.random.default_rng(int(null_frac * 100) + 3)
    page_rows = [64 * 300, 64 * 1100, 64 * 2, 64 * 120 + 37]
    page_bw = [3, 5, 6, 7]
    dst_bw = 7
    blob = bytearray()
    refs, want_rows, want_def = [], [], []
    for pi, (rows, bw) in enumerate(zip(page_rows, page_bw)):
        defs = rng.random(rows) >= null_frac if null_frac < 1.0 else np.zeros(rows, dtype=bool)
        if pi == 2 and null_frac > 0:
            defs[:] = False  # an all-NULL page: no width byte, no data (hdfs-parquet-scanner.cc:921)
        vals = rng.integers(0, 1 << bw, size=int(defs.sum())).astype(np.uint32)
        full = np.zeros(rows, dtype=np.uint32)
        full[defs] = vals
        want_rows.append(full)
        want_def.append(defs)
        dwords = oracle.pack(defs.astype(np.uint32), 1)
        blob += bytes(rng.integers(0, 255, size=int(rng.integers(17, 40)), dtype=np.uint8))
        blob += np.int32(dwords.nbytes).tobytes()
        def_off = len(blob)
        blob += dwords.tobytes()
        data_off, data_bytes = len(blob), 0
        if len(vals):
            words = oracle.pack(vals, bw)
            blob += bytes([bw])
            data_off, data_bytes = len(blob), words.nbytes
            blob += words.tobytes()
        refs.append(dict(src_off=data_off, src_bytes=data_bytes, src_bit_width=bw, groups=(rows + 63) // 64,
                         def_off=def_off, def_bytes=dwords.nbytes, rows=rows))
    blob += bytes(16)
    dblob = torch.from_numpy(np.frombuffer(bytes(blob), dtype=np.uint8).copy()).cuda()
    total_groups = sum(r["groups"] for r in refs)
    planes = torch.full((total_groups * dst_bw,), -1, dtype=torch.int64, device="cuda")
    dplanes = torch.full((total_groups,), -1, dtype=torch.int64, device="cuda")
    g0, crefs = 0, []
    for r in refs:
        crefs.append(dict(src_dev=dblob.data_ptr() + r["src_off"], src_bytes=r["src_bytes"], src_bit_width=r["src_bit_width"],
                          dst_bit_width=dst_bw, dst_dev=planes.data_ptr() + g0 * dst_bw * 8, groups=r["groups"],
                          def_src_dev=dblob.data_ptr() + r["def_off"], def_bytes=r["def_bytes"], rows=r["rows"], expand=1))
        # the definition levels themselves: a 1-bit page copied as it is
        crefs.append(dict(src_dev=dblob.data_ptr() + r["def_off"], src_bytes=r["def_bytes"], src_bit_width=1, dst_bit_width=1,
                          dst_dev=dplanes.data_ptr() + g0 * 8, groups=r["groups"], rows=r["rows"]))
        g0 += r["groups"]
    nulls = ctx.pages_gather_dev(crefs)
    torch.cuda.synchronize()
    assert int(nulls.item()) == sum(int((~d).sum()) for d in want_def)
    got = planes.cpu().numpy().view(np.uint64)
    gotd = dplanes.cpu().numpy().view(np.uint64)
    g0 = 0
    for r, full, defs in zip(refs, want_rows, want_def):
        dec = oracle.unpack(got[g0 * dst_bw:(g0 + r["groups"]) * dst_bw], dst_bw, 0, len(full))
        assert (dec == full).all()
        dd = oracle.unpack(gotd[g0:g0 + r["groups"]], 1, 0, len(full))
        assert (dd == defs.astype(np.uint32)).all()
        g0 += r["groups"]


def test_assemble_rows_null_indicators(ctx, oracle):
    """A selected row that is NULL in a materialised column gets its NULL indicator set and a zero slot
    (ReadSlot -> tuple->SetNull, hdfs-parquet-scanner.cc:515-531)."""
    import torch
    rng = np.random.default_rng(5)
    n = 64 * 3000 + 5
    sel = rng.random(n) < 0.05
    words = np.packbits(np.concatenate([sel, np.zeros((-n) % 64, dtype=bool)]), bitorder="little").view(np.uint64)
    bm = torch.from_numpy(words.view(np.int64).copy()).cuda()
    slots, keep, cols = [], [], []
    tuple_size = 16  # [null byte x 2][pad][int32 at 4][int64 at 8]
    for c, (ss, off, dt) in enumerate(((4, 4, np.int32), (8, 8, np.int64))):
        bw = 9 + c
        codes = rng.integers(0, 1 << bw, size=n).astype(np.uint32)
        defs = rng.random(n) >= (0.3 if c == 0 else 0.0)
        codes[~defs] = 0  # row-aligned planes as fle_pages_gather_dev (expand) leaves them
        d = rng.integers(-1000, 1000, size=1 << bw).astype(dt)
        planes = torch.from_numpy(oracle.pack(codes, bw).view(np.int64)).cuda()
        dplanes = torch.from_numpy(oracle.pack(defs.astype(np.uint32), 1).view(np.int64)).cuda()
        dd = torch.from_numpy(d).cuda()
        keep += [planes, dplanes, dd]
        slots.append(dict(planes_dev=planes.data_ptr(), bytes=planes.numel() * 8, bit_width=bw, slot_size=ss, dict_dev=dd.data_ptr(),
                          ndict=1 << bw, tuple_offset=off, def_planes_dev=dplanes.data_ptr() if c == 0 else None, null_offset=c))
        cols.append((codes, defs, d, off, ss))
    k = int(sel.sum())
    tuples, cnt, bad = ctx.assemble_rows_dev(bm, n, slots, tuple_size, k)
    torch.cuda.synchronize()
    assert int(cnt.item()) == k and int(bad.item()) == 0
    got = tuples.cpu().numpy()[:k]
    rows = np.nonzero(sel)[0]
    want = np.zeros((k, tuple_size), dtype=np.uint8)
    for c, (codes, defs, d, off, ss) in enumerate(cols):
        v = d[codes[rows]].copy()
        if c == 0:
            v[~defs[rows]] = 0
            want[:, c] = (~defs[rows]).astype(np.uint8)
        want[:, off:off + ss] = v.view(np.uint8).reshape(k, ss)
    assert (got == want).all()


@pytest.mark.parametrize("null_frac", [None, 0.0, 0.05, 0.7])
def test_pages_gather_places_ragged_pages(ctx, oracle, null_frac):
    """Pages of ANY row count (the reference's writer closes a dictionary page after 524225 rows,
    hdfs-parquet-table-writer.cc:387) are placed at their row offset of ONE row-aligned plane array: page k + 1
    continues in the group page k ended in (fle_page_ref.dst_bit).  null_frac None: required column (no definition
    levels); otherwise nullable pages, expanded through their definition levels, which travel as a 1-bit column."""
    import torch
    rng = np.random.default_rng(17 if null_frac is None else int(null_frac * 100) + 1)
    page_rows = [64 * 30 + 1, 524225, 17, 3, 64 * 2, 5000, 63, 1]
    page_bw = [3, 5, 6, 6, 7, 4, 2, 7]
    dst_bw = 7
    nullable = null_frac is not None
    blob = bytearray()
    refs, want_rows, want_def = [], [], []
    for pi, (rows, bw) in enumerate(zip(page_rows, page_bw)):
        defs = (rng.random(rows) >= null_frac) if nullable else np.ones(rows, dtype=bool)
        if nullable and null_frac > 0 and pi == 3:
            defs[:] = False
        vals = rng.integers(0, 1 << bw, size=int(defs.sum())).astype(np.uint32)
        full = np.zeros(rows, dtype=np.uint32)
        full[defs] = vals
        want_rows.append(full)
        want_def.append(defs)
        blob += bytes(rng.integers(0, 255, size=int(rng.integers(17, 40)), dtype=np.uint8))
        def_off = def_bytes = 0
        if nullable:
            dwords = oracle.pack(defs.astype(np.uint32), 1)
            blob += np.int32(dwords.nbytes).tobytes()
            def_off, def_bytes = len(blob), dwords.nbytes
            blob += dwords.tobytes()
        data_off, data_bytes = len(blob), 0
        if len(vals):
            words = oracle.pack(vals, bw)
            blob += bytes([bw])
            data_off, data_bytes = len(blob), words.nbytes
            blob += words.tobytes()
        refs.append(dict(src_off=data_off, src_bytes=data_bytes, bw=bw, def_off=def_off, def_bytes=def_bytes, rows=rows))
    blob += bytes(16)
    dblob = torch.from_numpy(np.frombuffer(bytes(blob), dtype=np.uint8).copy()).cuda()
    total = sum(page_rows)
    G = (total + 63) // 64
    planes = torch.full((G * dst_bw,), -1, dtype=torch.int64, device="cuda")
    dplanes = torch.full((G,), -1, dtype=torch.int64, device="cuda")
    row0, crefs = 0, []
    for r in refs:
        common = dict(groups=(r["rows"] + 63) // 64, rows=r["rows"], dst_bit=row0 % 64)
        ref = dict(src_dev=dblob.data_ptr() + r["src_off"], src_bytes=r["src_bytes"], src_bit_width=r["bw"], dst_bit_width=dst_bw,
                   dst_dev=planes.data_ptr() + (row0 // 64) * dst_bw * 8, **common)
        if nullable:
            ref.update(def_src_dev=dblob.data_ptr() + r["def_off"], def_bytes=r["def_bytes"], expand=1)
            crefs.append(dict(src_dev=dblob.data_ptr() + r["def_off"], src_bytes=r["def_bytes"], src_bit_width=1, dst_bit_width=1,
                              dst_dev=dplanes.data_ptr() + (row0 // 64) * 8, **common))
        crefs.append(ref)
        row0 += r["rows"]
    for rep in range(2):  # (twice: the shared groups are cleared by the call itself)
        nulls = ctx.pages_gather_dev(crefs)
    torch.cuda.synchronize()
    assert int(nulls.item()) == sum(int((~d).sum()) for d in want_def)
    got = planes.cpu().numpy().view(np.uint64)
    dec = oracle.unpack(got, dst_bw, 0, total)
    assert (dec == np.concatenate(want_rows)).all()
    if total % 64:  # rows past the end of the column read as 0
        assert (oracle.unpack(got, dst_bw, 0, G * 64)[total:] == 0).all()
    if nullable:
        dd = oracle.unpack(dplanes.cpu().numpy().view(np.uint64), 1, 0, total)
        assert (dd == np.concatenate(want_def).astype(np.uint32)).all()


@pytest.mark.parametrize("size", [1, 2, 4, 8])
@pytest.mark.parametrize("null_frac", [None, 0.0, 0.1])
def test_pages_gather_plain_pages_and_assemble(ctx, oracle, size, null_frac):
    """PLAIN pages (a column that overflowed its dictionary, hdfs-parquet-table-writer.cc:276-279): values back to back at
    any byte address, only the non-null ones for a nullable column, placed into a row-aligned value array; the assemble
    kernel copies a survivor's value into its slot (ParquetPlainEncoder::Decode) or sets its NULL indicator."""
    import torch
    dt = {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}[size]
    rng = np.random.default_rng(size * 10 + (0 if null_frac is None else int(null_frac * 100) + 1))
    page_rows = [64 * 30 + 1, 70001, 17, 64 * 2, 5000, 1]
    nullable = null_frac is not None
    blob = bytearray()
    refs, want_vals, want_def = [], [], []
    info = np.iinfo(dt)
    for rows in page_rows:
        defs = (rng.random(rows) >= null_frac) if nullable else np.ones(rows, dtype=bool)
        vals = rng.integers(info.min, info.max, size=int(defs.sum()), dtype=dt)
        full = np.zeros(rows, dtype=dt)
        full[defs] = vals
        want_vals.append(full)
        want_def.append(defs)
        blob += bytes(rng.integers(0, 255, size=int(rng.integers(17, 40)), dtype=np.uint8))
        def_off = def_bytes = 0
        if nullable:
            dwords = oracle.pack(defs.astype(np.uint32), 1)
            blob += np.int32(dwords.nbytes).tobytes()
            def_off, def_bytes = len(blob), dwords.nbytes
            blob += dwords.tobytes()
        refs.append(dict(src_off=len(blob), src_bytes=vals.nbytes, def_off=def_off, def_bytes=def_bytes, rows=rows))
        blob += vals.tobytes()
    blob += bytes(16)
    dblob = torch.from_numpy(np.frombuffer(bytes(blob), dtype=np.uint8).copy()).cuda()
    total = sum(page_rows)
    G = (total + 63) // 64
    values = torch.full((G * 64 * size,), -1, dtype=torch.int8, device="cuda")
    dplanes = torch.zeros(G, dtype=torch.int64, device="cuda")
    row0, crefs = 0, []
    for r in refs:
        ref = dict(src_dev=dblob.data_ptr() + r["src_off"], src_bytes=r["src_bytes"], plain_size=size,
                   dst_dev=values.data_ptr() + row0 * size, groups=(r["rows"] + 63) // 64, rows=r["rows"])
        if nullable:
            ref.update(def_src_dev=dblob.data_ptr() + r["def_off"], def_bytes=r["def_bytes"], expand=1)
            crefs.append(dict(src_dev=dblob.data_ptr() + r["def_off"], src_bytes=r["def_bytes"], src_bit_width=1, dst_bit_width=1,
                              dst_dev=dplanes.data_ptr() + (row0 // 64) * 8, groups=(r["rows"] + 63) // 64, rows=r["rows"], dst_bit=row0 % 64))
        crefs.append(ref)
        row0 += r["rows"]
    nulls = ctx.pages_gather_dev(crefs)
    torch.cuda.synchronize()
    want = np.concatenate(want_vals)
    defs = np.concatenate(want_def)
    assert int(nulls.item()) == int((~defs).sum())
    got = values.cpu().numpy().view(dt)[:total]
    assert (got == want).all()
    # survivors' tuples: [null byte][pad][slot]
    sel = rng.random(total) < 0.1
    words = np.packbits(np.concatenate([sel, np.zeros((-total) % 64, dtype=bool)]), bitorder="little").view(np.uint64)
    bm = torch.from_numpy(words.view(np.int64).copy()).cuda()
    tuple_size, off = 16, 8
    slot = dict(slot_size=size, tuple_offset=off, plain_values_dev=values.data_ptr(), bytes=values.numel(), null_offset=0,
                def_planes_dev=dplanes.data_ptr() if nullable else None)
    k = int(sel.sum())
    tuples, cnt, bad = ctx.assemble_rows_dev(bm, total, [slot], tuple_size, k)
    torch.cuda.synchronize()
    assert int(cnt.item()) == k and int(bad.item()) == 0
    rows = np.nonzero(sel)[0]
    wt = np.zeros((k, tuple_size), dtype=np.uint8)
    wt[:, 0] = (~defs[rows]).astype(np.uint8)
    wt[:, off:off + size] = want[rows].copy().view(np.uint8).reshape(k, size)
    assert (tuples.cpu().numpy()[:k] == wt).all()


@pytest.mark.parametrize("size", [4, 8])
@pytest.mark.parametrize("null_frac", [None, 0.2])
def test_pages_gather_mixed_dictionary_and_plain_chunk(ctx, oracle, size, null_frac):
    """A column chunk that starts dictionary-encoded and falls back to PLAIN after 40000 distinct values
    (hdfs-parquet-table-writer.cc:276-279): its dictionary pages are decoded through the dictionary into the same
    row-aligned value array the PLAIN pages are copied into."""
    import torch
    dt = {4: np.int32, 8: np.int64}[size]
    rng = np.random.default_rng(size + (0 if null_frac is None else 7))
    nullable = null_frac is not None
    d = rng.integers(-10**6, 10**6, size=1000).astype(dt)
    bw = 10
    page_rows = [64 * 40 + 3, 9000, 70001, 29]
    page_coded = [True, True, False, False]
    blob = bytearray()
    refs, want_vals, want_def = [], [], []
    for rows, coded in zip(page_rows, page_coded):
        defs = (rng.random(rows) >= null_frac) if nullable else np.ones(rows, dtype=bool)
        m = int(defs.sum())
        if coded:
            codes = rng.integers(0, len(d), size=m).astype(np.uint32)
            vals = d[codes]
        else:
            vals = rng.integers(-10**9, 10**9, size=m).astype(dt)
        full = np.zeros(rows, dtype=dt)
        full[defs] = vals
        want_vals.append(full)
        want_def.append(defs)
        blob += bytes(rng.integers(0, 255, size=int(rng.integers(17, 40)), dtype=np.uint8))
        def_off = def_bytes = 0
        if nullable:
            dwords = oracle.pack(defs.astype(np.uint32), 1)
            blob += np.int32(dwords.nbytes).tobytes()
            def_off, def_bytes = len(blob), dwords.nbytes
            blob += dwords.tobytes()
        if coded:
            words = oracle.pack(codes, bw) if m else np.zeros(0, dtype=np.uint64)
            blob += bytes([bw])
            refs.append(dict(src_off=len(blob), src_bytes=words.nbytes, def_off=def_off, def_bytes=def_bytes, rows=rows, coded=True))
            blob += words.tobytes()
        else:
            refs.append(dict(src_off=len(blob), src_bytes=vals.nbytes, def_off=def_off, def_bytes=def_bytes, rows=rows, coded=False))
            blob += vals.tobytes()
    blob += bytes(16)
    dblob = torch.from_numpy(np.frombuffer(bytes(blob), dtype=np.uint8).copy()).cuda()
    ddict = torch.from_numpy(d).cuda()
    total = sum(page_rows)
    G = (total + 63) // 64
    values = torch.full((G * 64 * size,), -1, dtype=torch.int8, device="cuda")
    row0, crefs = 0, []
    for r in refs:
        ref = dict(src_dev=dblob.data_ptr() + r["src_off"], src_bytes=r["src_bytes"], plain_size=size,
                   dst_dev=values.data_ptr() + row0 * size, groups=(r["rows"] + 63) // 64, rows=r["rows"])
        if r["coded"]:
            ref.update(src_bit_width=bw, dict_dev=ddict.data_ptr(), ndict=len(d))
        if nullable:
            ref.update(def_src_dev=dblob.data_ptr() + r["def_off"], def_bytes=r["def_bytes"], expand=1)
        crefs.append(ref)
        row0 += r["rows"]
    ctx.pages_gather_dev(crefs)
    torch.cuda.synchronize()
    got = values.cpu().numpy().view(dt)[:total]
    assert (got == np.concatenate(want_vals)).all()


def test_prefetch_marks(ctx, oracle):
    """fle_h2d_prefetch + fle_prefetch_mark / fle_prefetch_wait: two uploads on the copy stream, each evaluated behind its
    own mark (the second upload is enqueued BEFORE the first evaluation, as a scanner does); fence and error cases."""
    import torch
    import fle_b200
    from oracle.oracle import LT
    lib, h = ctx.lib, ctx.h
    rng = np.random.default_rng(21)
    n, bw = 64 * 3000 + 5, 9
    vals = [rng.integers(0, 1 << bw, size=n).astype(np.uint32) for _ in range(2)]
    words = [oracle.pack(v, bw) for v in vals]
    host = [torch.from_numpy(w.view(np.int64)).pin_memory() for w in words]
    dev = [torch.empty_like(hh, device="cuda") for hh in host]
    for k in range(2):
        assert lib.fle_h2d_prefetch(h, dev[k].data_ptr(), host[k].data_ptr(), host[k].numel() * 8) == 0
        assert lib.fle_prefetch_mark(h, k) == 0
    for k in range(2):
        assert lib.fle_prefetch_wait(h, k) == 0
        bm, cnt = ctx.eval_dev([(dev[k], bw)], [(LT, 0, 100)], n)
        torch.cuda.synchronize()
        assert int(cnt.item()) == int((vals[k] < 100).sum())
    assert lib.fle_prefetch_fence(h) == 0
    assert lib.fle_prefetch_wait(h, 7) == fle_b200.capi.E_INVALID   # never marked
    assert lib.fle_prefetch_mark(h, 8) == fle_b200.capi.E_INVALID   # slots 0..7
