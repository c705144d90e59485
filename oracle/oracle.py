"""TEST INFRASTRUCTURE ONLY: ctypes loaders for the two CPU checkers.

* ``Oracle``  -- oracle/libfle_oracle.so, the plain-C restatement (oracle/fle_oracle.c).
* ``Ref``     -- oracle/_ref/libfle_ref.so, the reference's own fle-encoding.h compiled
                 unmodified (oracle/ref_driver.cc); exists wherever oracle/Makefile ran with
                 /root/reference present, and travels to the GPU box as a prebuilt file.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

EQ, LT, LE, GT, GE, IN, AND, OR, ALL, NONE = range(10)
OP_NAMES = {EQ: "eq", LT: "lt", LE: "le", GT: "gt", GE: "ge", IN: "in"}

_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)


def _p(a, t):
    return a.ctypes.data_as(t)


def build(force: bool = False) -> None:
    """Compile the checkers (gcc/g++ via oracle/Makefile).  Building is not using."""
    if force:
        subprocess.check_call(["make", "-C", _HERE, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


def _lits(lits):
    a = np.ascontiguousarray(np.atleast_1d(np.asarray(lits, dtype=np.uint64)))
    return a


class _OrcDecoder(C.Structure):
    _fields_ = [("words", _u64p), ("ngroups_guard", C.c_int64), ("bw", C.c_int),
                ("count", C.c_int), ("next_group", C.c_int64), ("cur_group", C.c_int64)]


class Oracle:
    def __init__(self):
        path = os.path.join(_HERE, "libfle_oracle.so")
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path)
        L.orc_encoded_bytes.restype = C.c_int64
        L.orc_encoded_bytes.argtypes = [C.c_int64, C.c_int]
        L.orc_pack.restype = C.c_int64
        L.orc_pack.argtypes = [_u32p, C.c_int64, C.c_int, _u64p]
        L.orc_unpack.restype = None
        L.orc_unpack.argtypes = [_u64p, C.c_int, C.c_int64, C.c_int64, _u32p]
        for f in (L.orc_pred, L.orc_pred_planes):
            f.restype = C.c_int64
            f.argtypes = [C.c_int, _u64p, C.c_int, C.c_int64, C.c_int64, _u64p, C.c_int, _u64p]
        L.orc_intersect_bitset.restype = C.c_int64
        L.orc_intersect_bitset.argtypes = [_u64p, C.c_int64, _u64p]
        L.orc_decoder_init.restype = None
        L.orc_decoder_init.argtypes = [C.POINTER(_OrcDecoder), _u64p, C.c_int64, C.c_int]
        L.orc_decoder_get.restype = C.c_int
        L.orc_decoder_get.argtypes = [C.POINTER(_OrcDecoder), _u32p]
        L.orc_decoder_get_skip.restype = C.c_int
        L.orc_decoder_get_skip.argtypes = [C.POINTER(_OrcDecoder), _u32p, C.c_int]
        L.orc_decoder_skip.restype = C.c_int
        L.orc_decoder_skip.argtypes = [C.POINTER(_OrcDecoder), C.c_int]
        L.orc_decoder_pred.restype = None
        L.orc_decoder_pred.argtypes = [C.POINTER(_OrcDecoder), C.c_int, C.c_int64, _u64p, C.c_int,
                                       _u64p, _i64p]
        L.orc_dict_translate.restype = C.c_int
        L.orc_dict_translate.argtypes = [C.c_int, _i64p, C.c_int64, C.c_int64, _u64p]
        L.orc_dict_translate_in.restype = C.c_int64
        L.orc_dict_translate_in.argtypes = [_i64p, C.c_int64, _i64p, C.c_int64, _u64p]
        L.orc_dict_bit_width.restype = C.c_int
        L.orc_dict_bit_width.argtypes = [C.c_int64]

    # -- codec ----------------------------------------------------------------------
    def encoded_bytes(self, n, bw):
        return int(self.lib.orc_encoded_bytes(n, bw))

    def pack(self, vals, bw):
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
        out = np.zeros(self.encoded_bytes(len(vals), bw) // 8, dtype=np.uint64)
        self.lib.orc_pack(_p(vals, _u32p), len(vals), bw, _p(out, _u64p))
        return out

    def unpack(self, words, bw, first_row, n):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        out = np.zeros(n, dtype=np.uint32)
        self.lib.orc_unpack(_p(words, _u64p), bw, first_row, n, _p(out, _u32p))
        return out

    # -- predicates -----------------------------------------------------------------
    def pred(self, op, words, bw, first_row, num_rows, lits, planes=False):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        l = _lits(lits)
        out = np.zeros((num_rows + 63) // 64, dtype=np.uint64)
        f = self.lib.orc_pred_planes if planes else self.lib.orc_pred
        cnt = f(op, _p(words, _u64p), bw, first_row, num_rows, _p(l, _u64p), len(l), _p(out, _u64p))
        return out, int(cnt)

    def eval_tree(self, cols, postfix, n):
        """cols: list of (words, bw); postfix: list of (kind, col, lits) with kind in
        EQ..IN leaves, AND/OR inner nodes, ALL/NONE constants (simple-predicates.h:145-163)."""
        nw = (n + 63) // 64
        tail = np.uint64((1 << (n & 63)) - 1) if n & 63 else None
        st = []
        for kind, col, lits in postfix:
            if kind <= IN:
                w, bw = cols[col]
                st.append(self.pred(kind, w, bw, 0, n, lits)[0])
            elif kind == ALL:
                a = np.full(nw, np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
                if tail is not None:
                    a[-1] = tail
                st.append(a)
            elif kind == NONE:
                st.append(np.zeros(nw, dtype=np.uint64))
            else:
                b = st.pop()
                a = st.pop()
                st.append(a & b if kind == AND else a | b)
        assert len(st) == 1
        bm = st[0]
        return bm, int(np.unpackbits(bm.view(np.uint8)).sum())

    def intersect_bitset(self, root_words, n_root, sub_words):
        root = np.array(root_words, dtype=np.uint64, copy=True)
        sub = np.ascontiguousarray(sub_words, dtype=np.uint64)
        if len(sub) == 0:
            sub = np.zeros(1, dtype=np.uint64)
        cnt = self.lib.orc_intersect_bitset(_p(root, _u64p), n_root, _p(sub, _u64p))
        return root, int(cnt)

    # -- cursor model -----------------------------------------------------------------
    def decoder(self, words, nbytes, bw):
        return OracleDecoder(self, words, nbytes, bw)

    # -- dictionary -------------------------------------------------------------------
    def dict_translate(self, op, dict_sorted, literal):
        d = np.ascontiguousarray(dict_sorted, dtype=np.int64)
        code = C.c_uint64(0)
        k = self.lib.orc_dict_translate(op, _p(d, _i64p), len(d), int(literal), C.byref(code))
        return int(k), int(code.value)

    def dict_translate_in(self, dict_sorted, lits):
        d = np.ascontiguousarray(dict_sorted, dtype=np.int64)
        l = np.ascontiguousarray(lits, dtype=np.int64)
        codes = np.zeros(max(1, len(l)), dtype=np.uint64)
        k = self.lib.orc_dict_translate_in(_p(d, _i64p), len(d), _p(l, _i64p), len(l), _p(codes, _u64p))
        return codes[:k].copy()

    def dict_bit_width(self, n):
        return int(self.lib.orc_dict_bit_width(n))


class OracleDecoder:
    def __init__(self, orc, words, nbytes, bw):
        self.orc = orc
        self.words = np.ascontiguousarray(words, dtype=np.uint64)
        self.d = _OrcDecoder()
        orc.lib.orc_decoder_init(C.byref(self.d), _p(self.words, _u64p), nbytes, bw)

    def get(self):
        v = C.c_uint32(0)
        ok = self.orc.lib.orc_decoder_get(C.byref(self.d), C.byref(v))
        return bool(ok), int(v.value)

    def get_skip(self, skip):
        v = C.c_uint32(0)
        ok = self.orc.lib.orc_decoder_get_skip(C.byref(self.d), C.byref(v), skip)
        return bool(ok), int(v.value)

    def skip(self, skip):
        return bool(self.orc.lib.orc_decoder_skip(C.byref(self.d), skip))

    def pred(self, op, num_rows, lits):
        l = _lits(lits)
        out = np.zeros((num_rows + 63) // 64 + 1, dtype=np.uint64)
        nbits = C.c_int64(0)
        self.orc.lib.orc_decoder_pred(C.byref(self.d), op, num_rows, _p(l, _u64p), len(l),
                                      _p(out, _u64p), C.byref(nbits))
        assert nbits.value == num_rows
        return out[: (num_rows + 63) // 64]


class _RefNode(C.Structure):
    _fields_ = [("kind", C.c_int32), ("col", C.c_int32), ("lit_off", C.c_int32), ("nlits", C.c_int32)]


class _RefLeaf(C.Structure):
    _fields_ = [("op", C.c_int32), ("col", C.c_int32), ("lit", C.c_uint64)]


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libfle_ref.so"))


class Ref:
    """The compiled reference.  enc arguments are uint64 word arrays (as Oracle.pack makes)."""

    def __init__(self, lib_path: str | None = None):
        """lib_path: another build of oracle/ref_driver.cc -- tests compile the same driver against
        the product's drop-in fle-encoding.h to run identical call sequences on both."""
        path = lib_path or os.path.join(_HERE, "_ref", "libfle_ref.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle` where /root/reference is mounted")
        L = self.lib = C.CDLL(path)
        if not L.ref_has_avx2():
            raise RuntimeError("host CPU lacks AVX2; the reference cannot run here")
        L.ref_pack.restype = C.c_int64
        L.ref_pack.argtypes = [_u32p, C.c_int64, C.c_int, _u8p, C.c_int64]
        L.ref_unpack.restype = C.c_int64
        L.ref_unpack.argtypes = [_u8p, C.c_int64, C.c_int, C.c_int64, _u32p]
        L.ref_get_skip.restype = C.c_int64
        L.ref_get_skip.argtypes = [_u8p, C.c_int64, C.c_int, _i32p, C.c_int64, _u32p]
        L.ref_pred.restype = C.c_int64
        L.ref_pred.argtypes = [C.c_int, _u8p, C.c_int64, C.c_int, C.c_int64, C.c_int64, _u64p, C.c_int, _u64p]
        L.ref_pred_windowed.restype = C.c_int64
        L.ref_pred_windowed.argtypes = [C.c_int, _u8p, C.c_int64, C.c_int, C.c_int64, C.c_int, _u64p, C.c_int, _u64p]
        L.ref_eval_tree.restype = C.c_int64
        L.ref_eval_tree.argtypes = [C.POINTER(_u8p), _i64p, _i32p, C.c_int, C.POINTER(_RefNode), C.c_int,
                                    _u64p, C.c_int64, _u64p]
        L.ref_scan_materialize.restype = C.c_int64
        L.ref_scan_materialize.argtypes = [C.c_int, _u8p, C.c_int64, C.c_int, C.c_int64, C.c_int, _u64p,
                                           C.c_int, _u32p]
        L.ref_skip_seq.restype = C.c_int64
        L.ref_skip_seq.argtypes = [_u8p, C.c_int64, C.c_int, _i32p, C.c_int64, _i32p, _i32p]
        L.ref_set_compat.restype = None
        L.ref_set_compat.argtypes = [C.c_int]
        L.ref_scan_materialize_checked.restype = C.c_int64
        L.ref_scan_materialize_checked.argtypes = L.ref_scan_materialize.argtypes
        L.ref_time_and_leaves.restype = C.c_double
        L.ref_time_and_leaves.argtypes = [C.POINTER(_u8p), _i32p, C.c_int, C.POINTER(_RefLeaf), C.c_int,
                                          C.c_int64, C.c_int, C.c_int, C.c_int, _u64p]
        L.ref_time_pack.restype = C.c_double
        L.ref_time_pack.argtypes = [_u32p, C.c_int64, C.c_int, C.c_int, C.c_int]
        L.ref_time_unpack.restype = C.c_double
        L.ref_time_unpack.argtypes = [_u8p, C.c_int64, C.c_int, C.c_int, C.c_int, _u32p]

    def pack(self, vals, bw):
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
        g = max(1, (len(vals) + 63) // 64)
        cap = (g + 2) * bw * 8 + 64
        out = np.zeros(cap, dtype=np.uint8)
        n = self.lib.ref_pack(_p(vals, _u32p), len(vals), bw, _p(out, _u8p), cap - 64)
        if n < 0:
            raise RuntimeError("reference Put returned false")
        return out[:n].view(np.uint64).copy()

    def unpack(self, words, bw, n):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        out = np.zeros(n, dtype=np.uint32)
        got = self.lib.ref_unpack(_p(words, _u8p), words.nbytes, bw, n, _p(out, _u32p))
        return out, int(got)

    def get_skip(self, words, bw, skips):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        s = np.ascontiguousarray(skips, dtype=np.int32)
        out = np.zeros(len(s), dtype=np.uint32)
        got = self.lib.ref_get_skip(_p(words, _u8p), words.nbytes, bw, _p(s, _i32p), len(s), _p(out, _u32p))
        return out, int(got)

    def pred(self, op, words, bw, pre_get, num_rows, lits):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        l = _lits(lits)
        out = np.zeros((num_rows + 63) // 64 + 1, dtype=np.uint64)
        sz = self.lib.ref_pred(op, _p(words, _u8p), words.nbytes, bw, pre_get, num_rows, _p(l, _u64p),
                               len(l), _p(out, _u64p))
        assert sz == num_rows, (sz, num_rows)
        return out[: (num_rows + 63) // 64]

    def pred_windowed(self, op, words, bw, n, window, lits):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        l = _lits(lits)
        out = np.zeros((n + 63) // 64 + 1, dtype=np.uint64)
        sz = self.lib.ref_pred_windowed(op, _p(words, _u8p), words.nbytes, bw, n, window, _p(l, _u64p),
                                        len(l), _p(out, _u64p))
        assert sz == n
        return out[: (n + 63) // 64]

    def eval_tree(self, cols, postfix, n):
        """Same arguments as Oracle.eval_tree (ALL/NONE constants are not reference leaves)."""
        ws = [np.ascontiguousarray(w, dtype=np.uint64) for w, _ in cols]
        encs = (_u8p * len(ws))(*[_p(w, _u8p) for w in ws])
        lens = np.array([w.nbytes for w in ws], dtype=np.int64)
        bws = np.array([bw for _, bw in cols], dtype=np.int32)
        nodes = (_RefNode * len(postfix))()
        flat = []
        for i, (kind, col, lits) in enumerate(postfix):
            nodes[i].kind = kind
            nodes[i].col = col if col is not None else 0
            if kind <= IN:
                l = [int(x) for x in np.atleast_1d(lits)]
                nodes[i].lit_off = len(flat)
                nodes[i].nlits = len(l)
                flat += l
        fl = np.array(flat if flat else [0], dtype=np.uint64)
        out = np.zeros((n + 63) // 64 + 1, dtype=np.uint64)
        sz = self.lib.ref_eval_tree(encs, _p(lens, _i64p), _p(bws, _i32p), len(ws), nodes, len(postfix),
                                    _p(fl, _u64p), n, _p(out, _u64p))
        assert sz == n
        return out[: (n + 63) // 64]

    def scan_materialize(self, op, words, bw, n, window, lits):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        l = _lits(lits)
        out = np.zeros(n, dtype=np.uint32)
        k = self.lib.ref_scan_materialize(op, _p(words, _u8p), words.nbytes, bw, n, window, _p(l, _u64p),
                                          len(l), _p(out, _u32p))
        if k < 0:
            raise RuntimeError("reference Get(val, skip) returned false")
        return out[:k].copy()

    def skip_seq(self, words, nbytes, bw, skips):
        """[Skip(s) for s in skips] as 0/1, then whether a Get() still succeeds."""
        words = np.ascontiguousarray(words, dtype=np.uint64)
        s = np.ascontiguousarray(skips, dtype=np.int32)
        res = np.zeros(len(s), dtype=np.int32)
        fin = C.c_int32(0)
        self.lib.ref_skip_seq(_p(words, _u8p), nbytes, bw, _p(s, _i32p), len(s), _p(res, _i32p), C.byref(fin))
        return res.tolist(), int(fin.value)

    def set_compat(self, on: bool):
        self.lib.ref_set_compat(1 if on else 0)

    def scan_materialize_checked(self, op, words, nbytes, bw, n, window, lits):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        l = _lits(lits)
        out = np.zeros(n, dtype=np.uint32)
        k = self.lib.ref_scan_materialize_checked(op, _p(words, _u8p), nbytes, bw, n, window, _p(l, _u64p), len(l),
                                                  _p(out, _u32p))
        return out[:k].copy()

    # -- timing (bench.py cpu_baseline) ------------------------------------------------
    def time_and_leaves(self, cols, leaves, n, nthreads, mode=0, reps=1):
        """cols: list of (uint64 words padded by >= 8 words, bw); leaves: list of (op, col, lit).
        Returns (seconds for one pass, match_count)."""
        ws = [np.ascontiguousarray(w, dtype=np.uint64) for w, _ in cols]
        encs = (_u8p * len(ws))(*[_p(w, _u8p) for w in ws])
        bws = np.array([bw for _, bw in cols], dtype=np.int32)
        lv = (_RefLeaf * len(leaves))()
        for i, (op, col, lit) in enumerate(leaves):
            lv[i].op, lv[i].col, lv[i].lit = op, col, int(lit)
        cnt = C.c_uint64(0)
        s = self.lib.ref_time_and_leaves(encs, _p(bws, _i32p), len(ws), lv, len(leaves), n, nthreads,
                                         mode, reps, C.byref(cnt))
        return float(s), int(cnt.value)

    def time_pack(self, vals, bw, nthreads, reps=1):
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
        return float(self.lib.ref_time_pack(_p(vals, _u32p), len(vals), bw, nthreads, reps))

    def time_unpack(self, words_padded, n, bw, nthreads, reps=1):
        w = np.ascontiguousarray(words_padded, dtype=np.uint64)
        out = np.zeros(n, dtype=np.uint32)
        s = self.lib.ref_time_unpack(_p(w, _u8p), n, bw, nthreads, reps, _p(out, _u32p))
        return float(s), out
