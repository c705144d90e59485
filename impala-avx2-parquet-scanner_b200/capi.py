"""ctypes binding of the C ABI in include/fle_b200.h (libfle_b200.so).

This is the Python-side twin of the cgo/JNI-style stub INTEGRATION.md shows for the
reference's C++ tree: plain pointers and sizes only.  torch is used purely as a provider
of device memory and streams (tensor.data_ptr(), torch.cuda.current_stream()).

There is NO fallback: if the shared library is missing or no B200 is present, loading or
context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfle_b200.so")

# opcodes (fle_b200.h)
EQ, LT, LE, GT, GE, IN, AND, OR, ALL, NONE = range(10)
OK, E_INVALID, E_NOMEM, E_CUDA, E_BUFFER, E_ALIGN, E_UNSUPPORTED, E_NODEVICE, E_DICT_FULL = 0, -1, -2, -3, -4, -5, -6, -7, -8

EXPORTS = [
    "fle_ctx_create", "fle_ctx_destroy", "fle_ctx_set_stream", "fle_ctx_sync", "fle_ctx_set_option",
    "fle_ctx_get_option", "fle_ctx_launch_count", "fle_last_error", "fle_strerror", "fle_version",
    "fle_device_count", "fle_host_alloc", "fle_host_free", "fle_dev_alloc", "fle_dev_free", "fle_h2d",
    "fle_d2h", "fle_encoded_bytes", "fle_pack_u32", "fle_pack_u32_dev", "fle_unpack_u32",
    "fle_unpack_u32_dev", "fle_unpack_u8_dev", "fle_unpack_u16_dev", "fle_eval", "fle_eval_dev", "fle_eval_no_nulls_dev", "fle_bitmap_count_dev", "fle_intersect_bitset",
    "fle_intersect_bitset_dev", "fle_bitmap_select_dev", "fle_gather_dev", "fle_sum_product_dev",
    "fle_page_parse", "fle_realign_dev", "fle_plain_eval_dev", "fle_dict_encode_dev", "fle_eval_codegen",
    "fle_eval_sum_product_dev", "fle_group_create_local", "fle_group_export", "fle_group_open", "fle_group_close",
    "fle_h2d_prefetch", "fle_prefetch_fence", "fle_prefetch_mark", "fle_prefetch_wait", "fle_ctx_share_copy_stream", "fle_host_register", "fle_host_unregister", "fle_pages_gather_dev",
    "fle_assemble_rows_dev", "fle_group_status", "fle_group_flush", "fle_eval_group_dev", "fle_eval_sum_product_group_dev",
]
GROUP_HANDLE_BYTES = 64


class FleError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fle_b200 error {code}: {msg}")
        self.code = code


class ColDesc(C.Structure):
    _fields_ = [("planes", C.c_void_p), ("bytes", C.c_uint64), ("bit_width", C.c_int32),
                ("max_def_level", C.c_int32), ("def_planes", C.c_void_p), ("def_bytes", C.c_uint64)]


class PageLayout(C.Structure):
    _fields_ = [("def_off", C.c_uint32), ("def_bytes", C.c_uint32), ("data_off", C.c_uint32),
                ("data_bytes", C.c_uint32), ("bit_width", C.c_int32), ("reserved", C.c_int32)]


T_INT32, T_INT64, T_FLOAT, T_DOUBLE = range(4)


def group_create_local(ctxs: Sequence["FleContext"]):
    """Contexts of ONE process (same or different devices) whose totals are summed in-kernel."""
    lib = load_library()
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    rc = lib.fle_group_create_local(arr, len(ctxs))
    if rc != OK:
        raise FleError(rc, lib.fle_strerror(rc).decode())


class PageRef(C.Structure):
    _fields_ = [("src_dev", C.c_void_p), ("src_bytes", C.c_uint64), ("src_bit_width", C.c_int32), ("dst_bit_width", C.c_int32),
                ("dst_dev", C.c_void_p), ("groups", C.c_int64), ("def_src_dev", C.c_void_p), ("def_bytes", C.c_uint64),
                ("rows", C.c_int64), ("expand", C.c_int32), ("dst_bit", C.c_int32),
                ("plain_size", C.c_int32), ("ndict", C.c_int32), ("dict_dev", C.c_void_p)]


class SlotDesc(C.Structure):
    _fields_ = [("planes_dev", C.c_void_p), ("bytes", C.c_uint64), ("bit_width", C.c_int32), ("slot_size", C.c_int32),
                ("dict_dev", C.c_void_p), ("ndict", C.c_int32), ("tuple_offset", C.c_int32),
                ("def_planes_dev", C.c_void_p), ("null_offset", C.c_int32), ("reserved", C.c_int32),
                ("plain_values_dev", C.c_void_p)]


class PredNode(C.Structure):
    _fields_ = [("op", C.c_int32), ("col", C.c_int32), ("code", C.c_uint64),
                ("codes", C.POINTER(C.c_uint64)), ("ncodes", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def load_library() -> C.CDLL:
    """dlopen libfle_b200.so and declare every prototype.  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "fle_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64, u64p, i32 = C.c_void_p, C.c_int64, C.POINTER(C.c_uint64), C.c_int
    L.fle_ctx_create.argtypes = [i32, vp, C.POINTER(vp)]
    L.fle_ctx_destroy.argtypes = [vp]
    L.fle_ctx_set_stream.argtypes = [vp, vp]
    L.fle_ctx_sync.argtypes = [vp]
    L.fle_ctx_set_option.argtypes = [vp, C.c_char_p, i64]
    L.fle_ctx_get_option.argtypes = [vp, C.c_char_p]
    L.fle_ctx_get_option.restype = i64
    L.fle_ctx_launch_count.argtypes = [vp]
    L.fle_ctx_launch_count.restype = i64
    L.fle_last_error.argtypes = [vp]
    L.fle_last_error.restype = C.c_char_p
    L.fle_strerror.argtypes = [i32]
    L.fle_strerror.restype = C.c_char_p
    L.fle_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.fle_host_free.argtypes = [vp]
    L.fle_dev_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.fle_dev_free.argtypes = [vp, vp]
    L.fle_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.fle_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    L.fle_encoded_bytes.argtypes = [i64, i32]
    L.fle_encoded_bytes.restype = C.c_size_t
    for f in (L.fle_pack_u32, L.fle_pack_u32_dev):
        f.argtypes = [vp, vp, i64, i32, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    for f in (L.fle_unpack_u32, L.fle_unpack_u32_dev, L.fle_unpack_u8_dev, L.fle_unpack_u16_dev):
        f.argtypes = [vp, vp, C.c_size_t, i32, i64, i64, vp]
    for f in (L.fle_eval, L.fle_eval_dev):
        f.argtypes = [vp, C.POINTER(ColDesc), i32, C.POINTER(PredNode), i32, i64, i64, vp, vp]
    L.fle_eval_no_nulls_dev.argtypes = [vp, C.POINTER(ColDesc), i32, C.POINTER(PredNode), i32, i64, i64, vp, vp, vp]
    L.fle_bitmap_count_dev.argtypes = [vp, vp, i64, vp]
    for f in (L.fle_intersect_bitset, L.fle_intersect_bitset_dev):
        f.argtypes = [vp, vp, i64, vp, i64]
    L.fle_dict_encode_dev.argtypes = [vp, vp, i32, i64, i32, vp, C.POINTER(C.c_int), C.POINTER(C.c_int), vp, C.c_size_t,
                                      C.POINTER(C.c_size_t)]
    L.fle_page_parse.argtypes = [vp, C.c_size_t, i32, C.POINTER(PageLayout)]
    L.fle_realign_dev.argtypes = [vp, vp, C.c_size_t, vp]
    L.fle_plain_eval_dev.argtypes = [vp, vp, i32, i64, i32, vp, i32, i32, vp, vp]
    L.fle_bitmap_select_dev.argtypes = [vp, vp, i64, vp, i64, vp]
    L.fle_gather_dev.argtypes = [vp, vp, C.c_size_t, i32, vp, vp, i64, vp, i32, i32, vp]
    L.fle_sum_product_dev.argtypes = [vp, vp, i64, vp, vp, C.c_size_t, i32, vp, i32, vp, vp]
    L.fle_eval_sum_product_dev.argtypes = [vp, C.POINTER(ColDesc), i32, C.POINTER(PredNode), i32, i64, i64, vp, i32, vp, i32,
                                           vp, vp, vp]
    L.fle_eval_codegen.argtypes = [C.POINTER(ColDesc), i32, C.POINTER(PredNode), i32, i32, C.c_char_p, C.c_size_t,
                                   C.POINTER(C.c_size_t)]
    L.fle_h2d_prefetch.argtypes = [vp, vp, vp, C.c_size_t]
    L.fle_prefetch_fence.argtypes = [vp]
    L.fle_prefetch_mark.argtypes = [vp, i32]
    L.fle_prefetch_wait.argtypes = [vp, i32]
    L.fle_ctx_share_copy_stream.argtypes = [vp, vp]
    L.fle_host_register.argtypes = [vp, C.c_size_t]
    L.fle_host_unregister.argtypes = [vp]
    L.fle_pages_gather_dev.argtypes = [vp, C.POINTER(PageRef), i32, vp]
    L.fle_assemble_rows_dev.argtypes = [vp, vp, i64, C.POINTER(SlotDesc), i32, i32, vp, i64, vp, vp]
    L.fle_group_create_local.argtypes = [C.POINTER(vp), i32]
    L.fle_group_export.argtypes = [vp, vp]
    L.fle_group_open.argtypes = [vp, i32, i32, vp]
    L.fle_group_close.argtypes = [vp]
    L.fle_group_flush.argtypes = [vp]
    L.fle_group_status.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), u64p, u64p]
    L.fle_eval_group_dev.argtypes = L.fle_eval_dev.argtypes
    L.fle_eval_sum_product_group_dev.argtypes = L.fle_eval_sum_product_dev.argtypes
    _lib = L
    return L


def _build_program(program: Sequence[tuple]):
    """program: postfix list of (op, col, literal_or_list).  Returns (PredNode array, keepalive)."""
    nodes = (PredNode * len(program))()
    keep = []
    for i, item in enumerate(program):
        op = int(item[0])
        col = item[1] if len(item) > 1 and item[1] is not None else 0
        lit = item[2] if len(item) > 2 else None
        nodes[i].op = op
        nodes[i].col = int(col)
        if op == IN:
            arr = np.ascontiguousarray(np.atleast_1d(np.asarray(lit if lit is not None else [], dtype=np.uint64)))
            keep.append(arr)
            nodes[i].codes = arr.ctypes.data_as(C.POINTER(C.c_uint64)) if len(arr) else None
            nodes[i].ncodes = len(arr)
        elif op <= GE:
            nodes[i].code = int(np.atleast_1d(lit)[0]) if lit is not None else 0
    return nodes, keep


def eval_codegen(bit_widths: Sequence[int], program: Sequence[tuple], compile=False, agg_col: int = -1, skew: bool = False):
    """The run-time specialisation of a program's shape, no device needed (fle_eval_codegen).
    Returns (rc, text): the generated CUDA C++ (compile=False), the NVRTC log, '' on success
    (compile=True), or the cubin as bytes (compile=2)."""
    lib = load_library()
    cols = (ColDesc * len(bit_widths))()
    for i, bw in enumerate(bit_widths):
        cols[i].bit_width = int(bw)
    nodes, _keep = _build_program(program)
    cap = 1 << 20
    buf = C.create_string_buffer(cap)
    n = C.c_size_t(0)
    rc = lib.fle_eval_codegen(cols, len(bit_widths), nodes, len(nodes), int(compile) | ((agg_col + 1) << 8) | ((1 << 16) if skew else 0), buf, cap,
                              C.byref(n))
    if int(compile) == 2 and rc == OK:
        return rc, buf.raw[:n.value]
    return rc, buf.value.decode(errors="replace")


class FleContext:
    """One scanner thread's handle: a CUDA stream plus the library's device scratch."""

    def __init__(self, device: int = 0, stream: int | None = None):
        """stream: a cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream); 0 means the
        legacy default stream (passed to the library as cudaStreamLegacy); None lets the library
        create its own non-blocking stream."""
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.fle_ctx_create(device, self._stream_arg(stream), C.byref(h))
        if rc != OK:
            raise FleError(rc, self.lib.fle_strerror(rc).decode())
        self.h = h
        self.device = device

    @staticmethod
    def _stream_arg(stream):
        if stream is None:
            return None
        return C.c_void_p(stream if stream else 0x1)  # 0x1 == cudaStreamLegacy

    def close(self):
        if getattr(self, "h", None):
            self.lib.fle_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != OK:
            raise FleError(rc, f"{self.lib.fle_strerror(rc).decode()}: {self.lib.fle_last_error(self.h).decode()}")

    # -- context ---------------------------------------------------------------------
    def set_stream(self, stream: int | None):
        self._check(self.lib.fle_ctx_set_stream(self.h, self._stream_arg(stream)))

    def sync(self):
        self._check(self.lib.fle_ctx_sync(self.h))

    def set_option(self, name: str, value: int):
        self._check(self.lib.fle_ctx_set_option(self.h, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        return int(self.lib.fle_ctx_get_option(self.h, name.encode()))

    @property
    def launch_count(self) -> int:
        return int(self.lib.fle_ctx_launch_count(self.h))

    def encoded_bytes(self, n: int, bw: int) -> int:
        return int(self.lib.fle_encoded_bytes(n, bw))

    # -- host-buffer entry points (numpy) -------------------------------------------------
    def pack(self, vals: np.ndarray, bw: int) -> np.ndarray:
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
        out = np.empty(self.encoded_bytes(len(vals), bw) // 8, dtype=np.uint64)
        ln = C.c_size_t(0)
        self._check(self.lib.fle_pack_u32(self.h, vals.ctypes.data, len(vals), bw, out.ctypes.data, out.nbytes, C.byref(ln)))
        assert ln.value == out.nbytes
        return out

    def unpack(self, words: np.ndarray, bw: int, first_row: int, n: int) -> np.ndarray:
        words = np.ascontiguousarray(words)
        out = np.empty(n, dtype=np.uint32)
        self._check(self.lib.fle_unpack_u32(self.h, words.ctypes.data, words.nbytes, bw, first_row, n, out.ctypes.data))
        return out

    def eval(self, cols: Sequence, program: Sequence[tuple], n: int, first_row: int = 0):
        """cols: list of (words ndarray, bw) or (words, bw, def_words) for nullable columns.
        Returns (bitmap uint64 ndarray, match_count)."""
        descs = (ColDesc * len(cols))()
        keep = []
        for i, c in enumerate(cols):
            w = np.ascontiguousarray(c[0])
            keep.append(w)
            descs[i].planes = w.ctypes.data
            descs[i].bytes = w.nbytes
            descs[i].bit_width = int(c[1])
            if len(c) > 2 and c[2] is not None:
                d = np.ascontiguousarray(c[2])
                keep.append(d)
                descs[i].max_def_level = 1
                descs[i].def_planes = d.ctypes.data
                descs[i].def_bytes = d.nbytes
        nodes, keep2 = _build_program(program)
        bm = np.zeros((n + 63) // 64, dtype=np.uint64)
        cnt = C.c_uint64(0)
        self._check(self.lib.fle_eval(self.h, descs, len(cols), nodes, len(program), first_row, n,
                                      bm.ctypes.data if len(bm) else None, C.byref(cnt)))
        return bm, int(cnt.value)

    def intersect_bitset(self, root: np.ndarray, n_root: int, sub: np.ndarray, n_sub: int) -> np.ndarray:
        r = np.array(root, dtype=np.uint64, copy=True)
        s = np.ascontiguousarray(sub, dtype=np.uint64)
        self._check(self.lib.fle_intersect_bitset(self.h, r.ctypes.data, n_root, s.ctypes.data if len(s) else None, n_sub))
        return r

    # -- device entry points (torch tensors as memory) --------------------------------------
    def pack_dev(self, vals, bw: int, out=None):
        """vals: torch int32/uint32-compatible CUDA tensor (4-byte elements).  Returns an int64
        CUDA tensor of plane words."""
        import torch
        n = vals.numel()
        nbytes = self.encoded_bytes(n, bw)
        if out is None:
            out = torch.empty(nbytes // 8, dtype=torch.int64, device=vals.device)
        ln = C.c_size_t(0)
        self._check(self.lib.fle_pack_u32_dev(self.h, vals.data_ptr(), n, bw, out.data_ptr(), out.numel() * 8, C.byref(ln)))
        return out

    def unpack_dev(self, words, bw: int, first_row: int, n: int, out=None):
        import torch
        if out is None:
            out = torch.empty(n, dtype=torch.int32, device=words.device)
        self._check(self.lib.fle_unpack_u32_dev(self.h, words.data_ptr(), words.numel() * words.element_size(), bw,
                                                first_row, n, out.data_ptr()))
        return out

    def unpack_narrow_dev(self, words, bw: int, first_row: int, n: int, out=None):
        """1-byte values for bw <= 8, 2-byte values for bw <= 16 (FleDecoder::Get<T> for int8 / int16 columns)."""
        import torch
        es = 1 if bw <= 8 else 2
        if out is None:
            out = torch.empty(n, dtype=torch.uint8 if es == 1 else torch.int16, device=words.device)
        es = out.element_size()
        fn = self.lib.fle_unpack_u8_dev if es == 1 else self.lib.fle_unpack_u16_dev
        self._check(fn(self.h, words.data_ptr(), words.numel() * words.element_size(), bw, first_row, n, out.data_ptr()))
        return out

    def make_col_descs(self, cols: Sequence):
        descs = (ColDesc * len(cols))()
        for i, c in enumerate(cols):
            w = c[0]
            descs[i].planes = w.data_ptr()
            descs[i].bytes = w.numel() * w.element_size()
            descs[i].bit_width = int(c[1])
            if len(c) > 2 and c[2] is not None:
                d = c[2]
                descs[i].max_def_level = 1
                descs[i].def_planes = d.data_ptr()
                descs[i].def_bytes = d.numel() * d.element_size()
        return descs

    def eval_dev(self, cols: Sequence, program: Sequence[tuple], n: int, first_row: int = 0, bitmap=None, count=None,
                 prepared=None):
        """cols: list of (int64 CUDA tensor of plane words, bw[, def-level words]).  Enqueues on the
        context's stream; returns (bitmap int64 tensor, count int64 tensor[1])."""
        import torch
        dev = cols[0][0].device
        if bitmap is None:
            bitmap = torch.empty(max(1, (n + 63) // 64), dtype=torch.int64, device=dev)
        if count is None:
            count = torch.zeros(1, dtype=torch.int64, device=dev)
        if prepared is None:
            prepared = (self.make_col_descs(cols),) + _build_program(program)
        descs, nodes, _keep = prepared
        self._check(self.lib.fle_eval_dev(self.h, descs, len(cols), nodes, len(nodes), first_row, n,
                                          bitmap.data_ptr(), count.data_ptr()))
        return bitmap, count

    def eval_no_nulls_dev(self, cols: Sequence, program: Sequence[tuple], n: int, bitmap=None, out=None, prepared=None):
        """OPTIONAL columns, no host sync: returns (bitmap, out) with out = int64[2] CUDA tensor (count, NULL rows found);
        the result is valid iff out[1] == 0."""
        import torch
        dev = cols[0][0].device
        if bitmap is None:
            bitmap = torch.empty(max(1, (n + 63) // 64), dtype=torch.int64, device=dev)
        if out is None:
            out = torch.zeros(2, dtype=torch.int64, device=dev)
        if prepared is None:
            prepared = (self.make_col_descs(cols),) + _build_program(program)
        descs, nodes, _keep = prepared
        self._check(self.lib.fle_eval_no_nulls_dev(self.h, descs, len(cols), nodes, len(nodes), 0, n, bitmap.data_ptr(),
                                                   out.data_ptr(), out.data_ptr() + 8))
        return bitmap, out

    def bitmap_count_dev(self, bitmap, nbits: int, count=None):
        import torch
        if count is None:
            count = torch.zeros(1, dtype=torch.int64, device=bitmap.device)
        self._check(self.lib.fle_bitmap_count_dev(self.h, bitmap.data_ptr(), nbits, count.data_ptr()))
        return count

    def bitmap_select_dev(self, bitmap, nbits: int, cap: int | None = None, ids=None, count=None):
        """Ascending row ids of the set bits.  Returns (ids int32 tensor[cap], count int64 tensor[1])."""
        import torch
        cap = nbits if cap is None else cap
        if ids is None:
            ids = torch.empty(max(1, cap), dtype=torch.int32, device=bitmap.device)
        if count is None:
            count = torch.zeros(1, dtype=torch.int64, device=bitmap.device)
        self._check(self.lib.fle_bitmap_select_dev(self.h, bitmap.data_ptr(), nbits, ids.data_ptr(), cap, count.data_ptr()))
        return ids, count

    def gather_dev(self, words, bw: int, ids, count, cap: int, dict_values=None, out=None):
        """Codes (or dictionary values when dict_values is given) of the selected rows."""
        import torch
        if dict_values is None:
            if out is None:
                out = torch.empty(max(1, cap), dtype=torch.int32, device=words.device)
            dptr, nd, es = None, 0, 0
        else:
            if out is None:
                out = torch.empty(max(1, cap), dtype=dict_values.dtype, device=words.device)
            dptr, nd, es = dict_values.data_ptr(), dict_values.numel(), dict_values.element_size()
        self._check(self.lib.fle_gather_dev(self.h, words.data_ptr(), words.numel() * words.element_size(), bw,
                                            ids.data_ptr(), count.data_ptr(), cap, dptr, nd, es, out.data_ptr()))
        return out

    def sum_product_dev(self, bitmap, nrows: int, plain_i64, words, bw: int, dict_i64, out=None):
        """(sum, count) int64 tensors: SUM over selected rows of plain_i64[r] * dict_i64[code(r)]."""
        import torch
        if out is None:
            out = torch.zeros(2, dtype=torch.int64, device=bitmap.device)
        self._check(self.lib.fle_sum_product_dev(self.h, bitmap.data_ptr(), nrows, plain_i64.data_ptr(), words.data_ptr(),
                                                 words.numel() * words.element_size(), bw, dict_i64.data_ptr(),
                                                 dict_i64.numel(), out.data_ptr(), out.data_ptr() + 8))
        return out

    def eval_sum_product_dev(self, cols: Sequence, program: Sequence[tuple], n: int, plain_i64, code_col: int, dict_i64,
                             first_row: int = 0, bitmap=None, out=None, prepared=None):
        """Predicate program + SUM(plain_i64[r] * dict_i64[code of row r in cols[code_col]]) over the selected
        rows in one pass.  out: int64[2] CUDA tensor = (sum, count); bitmap: optional int64 tensor."""
        import torch
        if prepared is None:
            prepared = (self.make_col_descs(cols),) + _build_program(program)
        descs, nodes, _keep = prepared
        if out is None:
            out = torch.zeros(2, dtype=torch.int64, device=plain_i64.device)
        self._check(self.lib.fle_eval_sum_product_dev(
            self.h, descs, len(cols), nodes, len(nodes), first_row, n, plain_i64.data_ptr(), code_col, dict_i64.data_ptr(),
            dict_i64.numel(), bitmap.data_ptr() if bitmap is not None else None, out.data_ptr(), out.data_ptr() + 8))
        return out

    # -- groups of GPUs (totals summed inside the scan kernel) --------------------------------
    def group_export(self) -> bytes:
        buf = C.create_string_buffer(GROUP_HANDLE_BYTES)
        self._check(self.lib.fle_group_export(self.h, buf))
        return buf.raw

    def group_open(self, rank: int, world: int, handles: Sequence[bytes]):
        blob = b"".join(handles)
        assert len(blob) == world * GROUP_HANDLE_BYTES
        self._check(self.lib.fle_group_open(self.h, rank, world, blob))

    def group_flush(self):
        """Deferred mode: enqueue the collection of the last group call's totals."""
        self._check(self.lib.fle_group_flush(self.h))

    def group_close(self):
        self._check(self.lib.fle_group_close(self.h))

    def group_status(self):
        r, w, d, t = C.c_int(0), C.c_int(0), C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.fle_group_status(self.h, C.byref(r), C.byref(w), C.byref(d), C.byref(t)))
        return {"rank": r.value, "world": w.value, "reductions_done": d.value, "peers_timed_out": t.value}

    def eval_group_dev(self, cols: Sequence, program: Sequence[tuple], n: int, first_row: int = 0, bitmap=None, count=None,
                       prepared=None):
        """eval_dev over this GPU's rows; count = matches over all GPUs of the group."""
        import torch
        dev = cols[0][0].device
        if bitmap is None:
            bitmap = torch.empty(max(1, (n + 63) // 64), dtype=torch.int64, device=dev)
        if count is None:
            count = torch.zeros(1, dtype=torch.int64, device=dev)
        if prepared is None:
            prepared = (self.make_col_descs(cols),) + _build_program(program)
        descs, nodes, _keep = prepared
        self._check(self.lib.fle_eval_group_dev(self.h, descs, len(cols), nodes, len(nodes), first_row, n,
                                                bitmap.data_ptr(), count.data_ptr()))
        return bitmap, count

    def eval_sum_product_group_dev(self, cols: Sequence, program: Sequence[tuple], n: int, plain_i64, code_col: int, dict_i64,
                                   first_row: int = 0, bitmap=None, out=None, prepared=None):
        """eval_sum_product_dev over this GPU's rows; out = (sum, count) over all GPUs of the group."""
        import torch
        if prepared is None:
            prepared = (self.make_col_descs(cols),) + _build_program(program)
        descs, nodes, _keep = prepared
        if out is None:
            out = torch.zeros(2, dtype=torch.int64, device=plain_i64.device)
        self._check(self.lib.fle_eval_sum_product_group_dev(
            self.h, descs, len(cols), nodes, len(nodes), first_row, n, plain_i64.data_ptr(), code_col, dict_i64.data_ptr(),
            dict_i64.numel(), bitmap.data_ptr() if bitmap is not None else None, out.data_ptr(), out.data_ptr() + 8))
        return out

    # -- row-group batches ---------------------------------------------------------------------
    def pages_gather_dev(self, refs: Sequence[dict], null_rows=None):
        """refs: dicts with the fields of fle_page_ref (device addresses as ints)."""
        import torch
        arr = (PageRef * len(refs))()
        for i, r in enumerate(refs):
            for k, v in r.items():
                setattr(arr[i], k, v)
        if null_rows is None:
            null_rows = torch.zeros(1, dtype=torch.int64, device="cuda")
        self._check(self.lib.fle_pages_gather_dev(self.h, arr, len(refs), null_rows.data_ptr()))
        return null_rows

    def assemble_rows_dev(self, bitmap, nrows: int, slots: Sequence[dict], tuple_size: int, cap_rows: int):
        """slots: dicts with the fields of fle_slot_desc.  Returns (tuples uint8 [cap_rows, tuple_size], count, bad_codes)."""
        import torch
        arr = (SlotDesc * len(slots))()
        for i, r in enumerate(slots):
            for k, v in r.items():
                setattr(arr[i], k, v)
        tuples = torch.empty((max(1, cap_rows), tuple_size), dtype=torch.uint8, device=bitmap.device)
        out = torch.zeros(2, dtype=torch.int64, device=bitmap.device)
        self._check(self.lib.fle_assemble_rows_dev(self.h, bitmap.data_ptr(), nrows, arr, len(slots), tuple_size,
                                                   tuples.data_ptr(), cap_rows, out.data_ptr(), out.data_ptr() + 8))
        return tuples, out[0:1], out[1:2]

    def page_parse(self, page: np.ndarray, max_def_level: int) -> PageLayout:
        page = np.ascontiguousarray(page, dtype=np.uint8)
        out = PageLayout()
        self._check(self.lib.fle_page_parse(page.ctypes.data, page.nbytes, max_def_level, C.byref(out)))
        return out

    def realign_dev(self, blob, offset: int, nbytes: int, out=None):
        """Funnel-shift copy of blob[offset:offset+nbytes] (uint8 CUDA tensor) to an aligned tensor."""
        import torch
        if out is None:
            out = torch.empty((nbytes + 15) // 16 * 2, dtype=torch.int64, device=blob.device)
        self._check(self.lib.fle_realign_dev(self.h, blob.data_ptr() + offset, nbytes, out.data_ptr()))
        return out

    def plain_eval_dev(self, values, op: int, literals, reference_order: bool = False, bitmap=None, count=None):
        """values: int32/int64/float32/float64 CUDA tensor.  Returns (bitmap int64 tensor, count tensor)."""
        import torch
        et = {torch.int32: (T_INT32, np.int32), torch.int64: (T_INT64, np.int64), torch.float32: (T_FLOAT, np.float32),
              torch.float64: (T_DOUBLE, np.float64)}[values.dtype]
        n = values.numel()
        lits = np.ascontiguousarray(np.atleast_1d(np.asarray(literals, dtype=et[1])))
        if bitmap is None:
            bitmap = torch.empty(max(1, (n + 63) // 64), dtype=torch.int64, device=values.device)
        if count is None:
            count = torch.zeros(1, dtype=torch.int64, device=values.device)
        self._check(self.lib.fle_plain_eval_dev(self.h, values.data_ptr(), et[0], n, op, lits.ctypes.data if len(lits) else None,
                                                len(lits), 1 if reference_order else 0, bitmap.data_ptr(), count.data_ptr()))
        return bitmap, count

    def dict_encode_dev(self, values, max_entries: int = 40000):
        """values: int32/int64 CUDA tensor.  Returns (sorted dictionary int64 tensor[nd], bit_width,
        plane words int64 tensor).  Raises FleError(E_DICT_FULL) like DictEncoder::Put() == -1."""
        import torch
        et = T_INT32 if values.dtype == torch.int32 else T_INT64
        n = values.numel()
        d = torch.empty(max_entries, dtype=torch.int64, device=values.device)
        enc = torch.empty(max(1, self.encoded_bytes(n, 16) // 8), dtype=torch.int64, device=values.device)
        nd, bw, ln = C.c_int(0), C.c_int(0), C.c_size_t(0)
        self._check(self.lib.fle_dict_encode_dev(self.h, values.data_ptr(), et, n, max_entries, d.data_ptr(), C.byref(nd),
                                                 C.byref(bw), enc.data_ptr(), enc.numel() * 8, C.byref(ln)))
        return d[: nd.value], bw.value, enc[: ln.value // 8]

    def intersect_bitset_dev(self, root, n_root: int, sub, n_sub: int):
        self._check(self.lib.fle_intersect_bitset_dev(self.h, root.data_ptr(), n_root, sub.data_ptr(), n_sub))
        return root
