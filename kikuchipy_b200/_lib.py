"""ctypes binding of ``libkdi.so`` (the C ABI declared in ``include/kdi.h``).

There is deliberately no fallback: if the shared library has not been built, or no CUDA device
is present, the calls below raise.  Build with ``python -c "import __graft_entry__ as g;
g.build()"`` (or ``make -C kikuchipy_b200/csrc``).
"""

from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkdi.so")

KDI_OK = 0
KDI_EINVAL = -1
KDI_ECUDA = -2
KDI_ENOMEM = -3
KDI_EUNSUPPORTED = -4
KDI_EINTERNAL = -5

KDI_U8, KDI_U16, KDI_F32, KDI_F64 = 0, 1, 2, 3
KDI_NCC, KDI_NDP = 0, 1
KDI_HOST, KDI_DEVICE = 0, 1

OPT_COMPUTE_DTYPE = 0
OPT_CERT_SIGMAS = 1
OPT_FORCE_EXACT = 2
OPT_CTA_GROUP = 3
OPT_STRIP_TILES = 4
OPT_SUPERBLOCK = 5
OPT_L2_POLICY = 6
OPT_TILE_ROTATE = 7
OPT_MAX_STAGES = 8
OPT_OVERLAP = 9
OPT_SPLIT_SELECT = 10
OPT_GEMM_SMS = 11
OPT_DEP_FLAGS = 12
OPT_MIN_GROUPS = 13
OPT_POST_PER_GROUP = 14
OPT_GEMM_SERIAL = 15
OPT_SM_PARTITION = 16
OPT_POST_CORESIDENT = 17
OPT_BULK_NORMALIZE = 18
OPT_EARLY_SPLIT = 19
OPT_DIV_DOUBLE = 20
OPT_DICT_VIEW = 21
OPT_PROJECT_LIBM = 22
OPT_GEMM_DUAL = 23
OPT_CERT_STRICT = 24
OPT_CERT_WIDEN = 25

REFINE_ORI, REFINE_PC, REFINE_ORI_PC = 0, 1, 2

_DTYPES = {
    np.dtype(np.uint8): KDI_U8,
    np.dtype(np.uint16): KDI_U16,
    np.dtype(np.float32): KDI_F32,
    np.dtype(np.float64): KDI_F64,
}
_TORCH_DTYPES = {"torch.uint8": KDI_U8, "torch.float32": KDI_F32, "torch.float64": KDI_F64,
                 "torch.uint16": KDI_U16}


class RefineOptions(C.Structure):
    """``kdi_refine_options`` (include/kdi.h)."""

    _fields_ = [
        ("xatol", C.c_double),
        ("fatol", C.c_double),
        ("maxiter", C.c_int64),
        ("maxfev", C.c_int64),
        ("adaptive", C.c_int),
    ]


class Timings(C.Structure):
    _fields_ = [
        ("normalize_exp_ms", C.c_float),
        ("normalize_dict_ms", C.c_float),
        ("gemm_topk_ms", C.c_float),
        ("rescore_ms", C.c_float),
        ("fallback_ms", C.c_float),
        ("merge_ms", C.c_float),
        ("total_ms", C.c_float),
        ("gemm_launches", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("flagged_rows", C.c_int64),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
        ("model_rows", C.c_int64),
    ]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


# every symbol include/kdi.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _f32p = C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_float)
_i64p, _u8p = C.POINTER(C.c_int64), C.POINTER(C.c_uint8)
SIGNATURES = {
    "kdi_version": (_i, []),
    "kdi_init": (_i, [_i, C.POINTER(_vp)]),
    "kdi_destroy": (_i, [_vp]),
    "kdi_last_error": (C.c_char_p, [_vp]),
    "kdi_set_option": (_i, [_vp, _i, C.c_double]),
    "kdi_get_timings": (_i, [_vp, C.POINTER(Timings)]),
    "kdi_device_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), _i64p]),
    "kdi_stream": (_i, [_vp, C.POINTER(_vp)]),
    "kdi_host_alloc": (_i, [_vp, _i64, C.POINTER(_vp)]),
    "kdi_host_free": (_i, [_vp, _vp]),
    "kdi_set_signal_mask": (_i, [_vp, _vp, _i64]),
    "kdi_patterns_create": (_i, [_vp, _vp, _i, _i, _i64, _i64, _i, _vp, C.POINTER(_vp)]),
    "kdi_patterns_shape": (_i, [_vp, _i64p, _i64p]),
    "kdi_patterns_read": (_i, [_vp, _vp, _vp]),
    "kdi_patterns_destroy": (_i, [_vp, _vp]),
    "kdi_match_topk": (_i, [_vp, _vp, _vp, _i, _i64, _vp, _vp, _i]),
    "kdi_match_full": (_i, [_vp, _vp, _vp, _vp, _i]),
    "kdi_debug_gemm16": (_i, [_vp, _vp, _vp, _vp]),
    "kdi_refine_objective": (_i, [_vp, _vp, _i, _vp, _i, _i, _i64, _i, _i, _i, _vp, _i64, _vp, _i, _vp, _vp, _vp, _vp]),
    "kdi_scores_f64": (_i, [_vp, _vp, _i, _vp, _i64, _vp, _i, _i64, _i64, _i, _vp, _i, _vp]),
    "kdi_merge_topk": (_i, [_vp, _i64, _i, _i, _vp, _vp, _i, _vp, _vp, _i]),
    "kdi_dictionary_indexing": (
        _i,
        [_vp, _vp, _i, _i, _i64, _vp, _i, _i, _i64, _i64, _i, _i, _i64, _vp, _i64, _vp, _vp, _i],
    ),
    "kdi_job_begin": (_i, [_vp, _vp, _i, _i, _i64, _i64, _i64, _i, _i, _vp, _i64, C.POINTER(_vp)]),
    "kdi_job_append": (_i, [_vp, _vp, _vp, _i, _i, _i64]),
    "kdi_job_finish": (_i, [_vp, _vp, _vp, _vp, _i]),
    "kdi_job_abort": (_i, [_vp, _vp]),
    "kdi_candidate_capacity": (_i, [_i]),
    "kdi_candidate_capacity_ctx": (_i, [_vp, _i]),
    "kdi_certificate_bound": (C.c_double, [_i, _i64]),
    "kdi_shard_candidates": (
        _i,
        [_vp, _vp, _i, _i, _i64, _vp, _i, _i, _i64, _i64, _i, _i, _vp, _i64, _vp, _vp, C.POINTER(_vp)],
    ),
    "kdi_shard_rescore_owned": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "kdi_shard_finalize": (_i, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i, _i64, _vp, _vp, _vp, C.POINTER(_i)]),
    "kdi_shard_exact_rows": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "kdi_shard_release": (_i, [_vp, _vp]),
    "kdi_comm_bytes_needed": (_i64, [_i, _i64, _i, _i]),
    "kdi_comm_create": (_i, [_vp, _i, _i, _i64, C.POINTER(_vp), _vp]),
    "kdi_comm_connect": (_i, [_vp, _vp, _vp]),
    "kdi_comm_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), _i64p]),
    "kdi_comm_destroy": (_i, [_vp, _vp]),
    "kdi_shard_run_peer": (
        _i, [_vp, _vp, _vp, _i, _i, _i64, _vp, _i, _i, _i64, _i64, _i, _i, _vp, _i64, _vp, _vp, _vp,
             C.POINTER(_i), C.POINTER(_vp)]),
    "kdi_shard_run_peer_projected": (
        _i, [_vp, _vp, _vp, _i, _i, _i64, _i64, _vp, _vp, _i, _i64, _i, _i, _vp, _i64, _vp, _vp, _vp,
             C.POINTER(_i), C.POINTER(_vp)]),
    "kdi_master_pattern_create": (_i, [_vp, _vp, _vp, _i, _i64, _i64, _vp, _i64, C.c_double, _i, C.c_double,
                                       C.c_double, C.POINTER(_vp)]),
    "kdi_master_pattern_destroy": (_i, [_vp, _vp]),
    "kdi_project_patterns": (_i, [_vp, _vp, _vp, _i, _i64, _vp, _i]),
    "kdi_project_patterns_varying_pc": (_i, [_vp, _vp, _vp, _i64, _vp, _i, _i, _vp, _vp]),
    "kdi_patterns_create_projected": (_i, [_vp, _vp, _vp, _i, _i64, _i, C.POINTER(_vp)]),
    "kdi_dictionary_indexing_projected": (
        _i, [_vp, _vp, _i, _i, _i64, _i64, _vp, _vp, _i, _i64, _i, _i, _vp, _i64, _vp, _vp, _i]),
    "kdi_shard_candidates_projected": (
        _i, [_vp, _vp, _i, _i, _i64, _i64, _vp, _vp, _i, _i64, _i, _i, _vp, _i64, _vp, _vp, C.POINTER(_vp)]),
    "kdi_orientation_similarity_map": (
        _i,
        [_vp, _vp, _i64, _i64, _i, _i, _i, _i, _vp, _i, _i, _i, _vp],
    ),
    "kdi_preprocess_patterns": (
        _i, [_vp, _vp, _i, _i, _i64, _i, _i, _i, _vp, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i]),
    "kdi_average_neighbour_patterns": (
        _i, [_vp, _vp, _i, _i, _i64, _i64, _i64, _vp, _i, _i, _vp, _vp, _i]),
    "kdi_refine": (
        _i,
        [_vp, _vp, _i, _vp, _i, _i, _i64, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp,
         C.POINTER(RefineOptions), _vp],
    ),
    "kdi_merge_crystal_maps": (
        _i,
        [_vp, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    ),
}

_lib = None
_lib_lock = threading.Lock()


def load():
    """Load ``libkdi.so`` and declare its signatures.  Raises if it is missing."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} has not been built; there is no CPU fallback. Build it with "
                "`make -C kikuchipy_b200/csrc` or `__graft_entry__.build()`."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


class KdiError(RuntimeError):
    """CUDA / internal failure reported by libkdi."""


def _raise(code: int, msg: str):
    if code == KDI_EINVAL:
        raise ValueError(msg)
    if code == KDI_ENOMEM:
        raise MemoryError(msg)
    if code == KDI_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise KdiError(f"libkdi error {code}: {msg}")


def _buffer(x, ctx=None):
    """(pointer, location, dtype code, keep-alive object) of a NumPy array or CUDA torch tensor."""
    if hasattr(x, "data_ptr") and hasattr(x, "is_cuda"):  # torch tensor
        if not x.is_contiguous():
            x = x.contiguous()
        code = _TORCH_DTYPES.get(str(x.dtype))
        if code is None:
            raise ValueError(f"unsupported tensor dtype {x.dtype}")
        if x.is_cuda:
            if ctx is not None:
                ctx._stream_sync(x.device)
            else:
                import torch

                torch.cuda.current_stream(x.device).synchronize()
            return x.data_ptr(), KDI_DEVICE, code, x
        return x.data_ptr(), KDI_HOST, code, x
    a = np.asarray(x)
    if a.dtype not in _DTYPES:
        a = a.astype(np.float32)  # the reference casts to the metric dtype first anyway
    a = np.ascontiguousarray(a)
    return a.ctypes.data, KDI_HOST, _DTYPES[a.dtype], a


class Patterns:
    """Device-resident prepared (normalised) pattern set: what ``prepare_*`` returns."""

    def __init__(self, ctx: "Context", handle: int):
        self._ctx = ctx
        self._h = handle
        rows, s_eff = C.c_int64(), C.c_int64()
        ctx._lib.kdi_patterns_shape(handle, C.byref(rows), C.byref(s_eff))
        self.shape = (rows.value, s_eff.value)

    def __array__(self, dtype=None, copy=None):
        out = np.empty(self.shape, dtype=np.float32)
        self._ctx._check(self._ctx._lib.kdi_patterns_read(self._ctx._h, self._h, out.ctypes.data))
        return out if dtype is None else out.astype(dtype)

    def compute(self):
        return np.asarray(self)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        if self._h:
            self._ctx._lib.kdi_patterns_destroy(self._ctx._h, self._h)
            self._h = None

    def __del__(self):
        try:
            if self._ctx._h:
                self.close()
        except Exception:
            pass


class Shard:
    """This rank's prepared experimental set + dictionary shard between the steps of the
    sharded pipeline (``kdi_shard_*`` in include/kdi.h)."""

    def __init__(self, ctx: "Context", handle: int, rows: int, kc: int, keep=None):
        self._ctx, self._h, self.rows, self.kc = ctx, handle, rows, kc
        # a device-resident float32 dictionary may be held as a VIEW of the caller's rows
        # (KDI_OPT_DICT_VIEW): the buffer has to outlive the shard
        self._keep = keep

    def rescore_owned(self, gidx, approx=None, keep_n: int = 0):
        """Exact scores of the candidates in ``gidx`` (at least ``rows`` x kc) whose dictionary rows
        this rank holds; -inf elsewhere (also in rows past ``rows``: padding of the row split).
        With ``approx`` (the merged tensor-core scores) and ``keep_n``, candidates that cannot
        reach the top ``keep_n`` are not read (``finalize`` verifies that)."""
        import torch

        n = max(int(gidx.shape[0]), self.rows)
        exact = torch.empty((n, self.kc), dtype=torch.float32, device=gidx.device)
        if n > self.rows:
            exact[self.rows:].fill_(-float("inf"))
        self._ctx._stream_sync(gidx.device)
        self._ctx._check(self._ctx._lib.kdi_shard_rescore_owned(
            self._ctx._h, self._h, gidx.data_ptr(), approx.data_ptr() if approx is not None else None,
            int(keep_n), exact.data_ptr()))
        return exact

    def finalize(self, approx, gidx, exact, keep_n: int, dict_total: int, row0: int = 0, rows: int | None = None):
        """Rank + certificate for the experimental rows ``[row0, row0 + rows)``; the three list
        tensors address that slice.  Returns ``(indices, scores, flagged row numbers)``."""
        import torch

        dev = gidx.device
        rows = self.rows - row0 if rows is None else int(rows)
        n_out = max(rows, int(gidx.shape[0]))  # callers may pass a padded slice
        scores = torch.empty((n_out, keep_n), dtype=torch.float32, device=dev)
        idx = torch.empty((n_out, keep_n), dtype=torch.int64, device=dev)
        if n_out > rows:
            scores[rows:].fill_(-float("inf"))
            idx[rows:].fill_(-1)
        flags = torch.empty((max(rows, 1),), dtype=torch.int32, device=dev)
        n_flag = C.c_int(0)
        self._ctx._stream_sync(dev)
        self._ctx._check(
            self._ctx._lib.kdi_shard_finalize(
                self._ctx._h, self._h, int(row0), int(rows), approx.data_ptr(), gidx.data_ptr(), exact.data_ptr(),
                int(keep_n), int(dict_total), scores.data_ptr(), idx.data_ptr(), flags.data_ptr(), C.byref(n_flag),
            )
        )
        return idx, scores, flags[: n_flag.value]

    def exact_rows(self, rows, keep_n: int):
        import torch

        n = int(rows.numel())
        scores = torch.empty((n, keep_n), dtype=torch.float32, device=rows.device)
        idx = torch.empty((n, keep_n), dtype=torch.int64, device=rows.device)
        rows = rows.to(torch.int32).contiguous()
        self._ctx._stream_sync(rows.device)
        self._ctx._check(
            self._ctx._lib.kdi_shard_exact_rows(self._ctx._h, self._h, rows.data_ptr(), n, int(keep_n),
                                                scores.data_ptr(), idx.data_ptr())
        )
        return idx, scores

    def close(self):
        if self._h:
            self._ctx._lib.kdi_shard_release(self._ctx._h, self._h)
            self._h = None
            self._keep = None

    def __del__(self):
        try:
            if self._ctx._h:
                self.close()
        except Exception:
            pass


IPC_HANDLE_BYTES = 64


class PeerComm:
    """This rank's symmetric block + the mapped blocks of the other ranks (``kdi_comm`` in
    include/kdi.h): the memory the sharded pipeline exchanges its lists through."""

    def __init__(self, ctx: "Context", rank: int, world: int, nbytes: int):
        self._ctx, self.rank, self.world, self.nbytes = ctx, rank, world, int(nbytes)
        h = _vp()
        self.handle = (C.c_uint8 * IPC_HANDLE_BYTES)()
        ctx._check(ctx._lib.kdi_comm_create(ctx._h, rank, world, int(nbytes), C.byref(h), self.handle))
        self._h = h.value

    def connect(self, handles: bytes):
        """``handles``: the ``world`` handles (64 bytes each) in rank order."""
        if len(handles) != self.world * IPC_HANDLE_BYTES:
            raise ValueError("one IPC handle per rank is needed")
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        self._ctx._check(self._ctx._lib.kdi_comm_connect(self._ctx._h, self._h, buf))

    def close(self):
        if self._h:
            self._ctx._lib.kdi_comm_destroy(self._ctx._h, self._h)
            self._h = None

    def __del__(self):
        try:
            if self._ctx._h:
                self.close()
        except Exception:
            pass


class IndexingJob:
    """An appendable dictionary-indexing job (``kdi_job_*`` in include/kdi.h)."""

    def __init__(self, ctx: "Context", handle: int, rows: int, keep_n: int, S: int, dict_rows: int):
        self._ctx, self._h, self.rows, self.keep_n, self.S, self.dict_rows = ctx, handle, rows, keep_n, S, dict_rows
        self.appended = 0

    def append(self, chunk):
        """The next dictionary rows (NumPy array or CUDA tensor, ``(n, ...)`` with ``S`` values per
        row).  The buffer may be reused as soon as the call returns."""
        ptr, loc, code, keep = _buffer(chunk, self._ctx)
        n = int(np.prod(chunk.shape))
        if n % self.S:
            raise ValueError(f"chunk of {n} values is not a whole number of {self.S}-pixel patterns")
        try:
            self._ctx._check(self._ctx._lib.kdi_job_append(self._ctx._h, self._h, ptr, loc, code, n // self.S))
        except Exception:
            self.abort()
            raise
        self.appended += n // self.S
        del keep

    def finish(self, out=None):
        """``(indices, scores)``: NumPy arrays, or the CUDA tensors given in ``out``."""
        if out is None:
            scores = np.empty((self.rows, self.keep_n), dtype=np.float32)
            idx = np.empty((self.rows, self.keep_n), dtype=np.int64)
            sptr, iptr, oloc = scores.ctypes.data, idx.ctypes.data, KDI_HOST
        else:
            idx, scores = out
            if hasattr(scores, "is_cuda") and scores.is_cuda:
                sptr, iptr, oloc = scores.data_ptr(), idx.data_ptr(), KDI_DEVICE
            else:
                sptr, iptr, oloc = scores.ctypes.data, idx.ctypes.data, KDI_HOST
        h, self._h = self._h, None  # the library frees the job whatever happens
        self._ctx._check(self._ctx._lib.kdi_job_finish(self._ctx._h, h, sptr, iptr, oloc))
        return idx, scores

    def abort(self):
        if self._h:
            self._ctx._lib.kdi_job_abort(self._ctx._h, self._h)
            self._h = None

    def __del__(self):
        try:
            if self._ctx._h:
                self.abort()
        except Exception:
            pass


class MasterPattern:
    """Device-resident master pattern + detector direction cosines (``kdi_master_pattern``)."""

    def __init__(self, ctx: "Context", handle: int, n_pixels: int):
        self._ctx, self._h, self.n_pixels = ctx, handle, n_pixels

    def close(self):
        if self._h:
            self._ctx._lib.kdi_master_pattern_destroy(self._ctx._h, self._h)
            self._h = None

    def __del__(self):
        try:
            if self._ctx._h:
                self.close()
        except Exception:
            pass


def _rotations(rot):
    """(pointer, location, keep-alive, n) of (n, 4) float64 quaternions (NumPy or CUDA tensor)."""
    if hasattr(rot, "is_cuda") and rot.is_cuda:
        import torch

        if rot.dtype != torch.float64 or not rot.is_contiguous():
            rot = rot.to(torch.float64).contiguous()
        return rot.data_ptr(), KDI_DEVICE, rot, int(rot.shape[0])
    r = np.ascontiguousarray(np.asarray(rot, dtype=np.float64).reshape(-1, 4))
    return r.ctypes.data, KDI_HOST, r, int(r.shape[0])


class Context:
    """One libkdi context = one (process, device)."""

    def __init__(self, device: int = 0):
        self._lib = load()
        h = _vp()
        rc = self._lib.kdi_init(device, C.byref(h))
        if rc != KDI_OK:
            _raise(rc, (self._lib.kdi_last_error(None) or b"").decode())
        self._h = h.value
        self.device = device
        self._signal_mask_key = None
        self._signal_mask_set = False

    # -- plumbing -------------------------------------------------------------
    def _check(self, rc: int):
        if rc != KDI_OK:
            _raise(rc, (self._lib.kdi_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            self.__dict__.pop("_pinned", None)  # kdi_destroy frees every pinned block still alive
            self._lib.kdi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream_sync(self, device):
        """Order work queued on torch's current stream before library calls (which run on the
        context's own stream).  Nothing to do when the current stream IS the context's stream."""
        import torch

        cur = torch.cuda.current_stream(device)
        if cur.cuda_stream != self.stream_handle():
            cur.synchronize()

    def torch_stream(self):
        """The context's stream as a ``torch.cuda.ExternalStream`` (make it current to avoid
        host synchronisation between torch ops / NCCL collectives and library calls)."""
        import torch

        return torch.cuda.ExternalStream(self.stream_handle(), device=torch.device("cuda", self.device))

    def set_option(self, option: int, value: float):
        self._check(self._lib.kdi_set_option(self._h, option, float(value)))

    def timings(self) -> dict:
        t = Timings()
        self._check(self._lib.kdi_get_timings(self._h, C.byref(t)))
        return t.as_dict()

    def device_info(self) -> dict:
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        self._check(self._lib.kdi_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "total_mem": mem.value}

    def stream_handle(self) -> int:
        s = _vp()
        self._check(self._lib.kdi_stream(self._h, C.byref(s)))
        return s.value or 0

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """NumPy array backed by pinned host memory.  Release it with :meth:`pinned_free` as soon as
        it is no longer needed (page-locked memory is not swappable); whatever is still alive is
        freed with the context (``kdi_destroy``)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = _vp()
        self._check(self._lib.kdi_host_alloc(self._h, n, C.byref(p)))
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self.__dict__.setdefault("_pinned", {})[p.value] = buf
        return arr

    def pinned_free(self, arr: np.ndarray) -> None:
        """Free the pinned block behind an array returned by :meth:`pinned_empty` (the array must
        not be used afterwards)."""
        addr = arr.__array_interface__["data"][0] if arr.size else None
        pinned = self.__dict__.get("_pinned", {})
        if addr is None or addr not in pinned:
            raise ValueError("not the start of a block returned by pinned_empty() of this context")
        del pinned[addr]
        self._check(self._lib.kdi_host_free(self._h, addr))

    def to_device(self, array: np.ndarray):
        """Upload a host array to a CUDA tensor through a small reusable pinned ring (two 32 MB
        blocks owned by the context) instead of one dataset-sized page-locked allocation: the host
        copy into one block overlaps the DMA out of the other."""
        import torch

        a = np.ascontiguousarray(array)
        dev = torch.device("cuda", self.device)
        out = torch.empty(a.shape, dtype=torch.from_numpy(a.reshape(-1)[:0]).dtype, device=dev)
        src = a.reshape(-1).view(np.uint8)
        dst = out.reshape(-1).view(torch.uint8)
        ring = self.__dict__.get("_ring")
        if ring is None:
            blk = 32 << 20
            ring = self.__dict__["_ring"] = [(self.pinned_empty((blk,), np.uint8), torch.cuda.Event()) for _ in range(2)]
        blk = ring[0][0].size
        stream = torch.cuda.current_stream(dev)
        for i, a0 in enumerate(range(0, src.size, blk)):
            buf, ev = ring[i & 1]
            if i >= 2:
                ev.synchronize()  # the DMA that last read this block has finished
            n = min(blk, src.size - a0)
            buf[:n] = src[a0:a0 + n]
            dst[a0:a0 + n].copy_(torch.from_numpy(buf[:n]), non_blocking=True)
            ev.record(stream)
        stream.synchronize()
        return out

    # -- masks and pattern sets ---------------------------------------------------
    def set_signal_mask(self, mask):
        """Signal mask for pattern sets created afterwards (True = pixel excluded).  A mask equal to
        the current one is not uploaded again."""
        key = None if mask is None else np.ascontiguousarray(np.asarray(mask).ravel().astype(np.uint8)).tobytes()
        if key == self._signal_mask_key and (key is not None or self._signal_mask_set):
            return
        if mask is None:
            self._check(self._lib.kdi_set_signal_mask(self._h, None, 0))
        else:
            m = np.frombuffer(key, dtype=np.uint8)
            self._check(self._lib.kdi_set_signal_mask(self._h, m.ctypes.data, m.size))
        self._signal_mask_key, self._signal_mask_set = key, True

    def patterns(self, data, rows: int, metric: int, row_mask=None) -> Patterns:
        """cast -> reshape (rows, -1) -> row mask -> signal mask -> normalise, on the device."""
        ptr, loc, code, keep = _buffer(data, self)
        n = int(np.prod(data.shape))
        if rows < 1 or n % rows:
            raise ValueError(f"cannot reshape array of size {n} into ({rows}, -1)")
        S = n // rows
        rm = None
        if row_mask is not None:
            rm = np.ascontiguousarray(np.asarray(row_mask).ravel().astype(np.uint8))
            if rm.size != rows:
                raise ValueError("navigation mask size does not match the number of patterns")
        h = _vp()
        self._check(
            self._lib.kdi_patterns_create(
                self._h, ptr, loc, code, rows, S, metric, rm.ctypes.data if rm is not None else None, C.byref(h)
            )
        )
        del keep
        return Patterns(self, h.value)

    # -- matching ------------------------------------------------------------------
    def match_topk(self, exp: Patterns, dic: Patterns, keep_n: int, index_offset: int = 0):
        m = exp.shape[0]
        scores = np.empty((m, keep_n), dtype=np.float32)
        idx = np.empty((m, keep_n), dtype=np.int64)
        self._check(
            self._lib.kdi_match_topk(
                self._h, exp._h, dic._h, keep_n, index_offset, scores.ctypes.data, idx.ctypes.data, KDI_HOST
            )
        )
        return idx, scores

    def match_full(self, exp: Patterns, dic: Patterns) -> np.ndarray:
        out = np.empty((exp.shape[0], dic.shape[0]), dtype=np.float32)
        self._check(self._lib.kdi_match_full(self._h, exp._h, dic._h, out.ctypes.data, KDI_HOST))
        return out

    def debug_gemm16(self, exp: Patterns, dic: Patterns) -> np.ndarray:
        out = np.empty((exp.shape[0], dic.shape[0]), dtype=np.float32)
        self._check(self._lib.kdi_debug_gemm16(self._h, exp._h, dic._h, out.ctypes.data))
        return out

    def device_rows(self, data, rows: int):
        """``data`` reshaped to ``(rows, -1)`` as a CUDA tensor (uploaded through the pinned ring when it
        lives on the host; 16-bit unsigned data travels as its bytes)."""
        import torch

        if hasattr(data, "is_cuda"):
            t = data if data.is_cuda else data.to(torch.device("cuda", self.device))
            return t.contiguous().reshape(int(rows), -1)
        a = np.asarray(data)
        if a.dtype not in _DTYPES:
            a = a.astype(np.float32)
        return self.to_device(np.ascontiguousarray(a).reshape(int(rows), -1))

    def scores_f64(self, exp_rows_dev, exp_row_index, dict_rows_dev, metric: int, candidates) -> np.ndarray:
        """float64 NCC / NDP scores of every row's listed candidates from the RAW device-resident patterns
        (``kdi_scores_f64``; the context's signal mask applies).  ``exp_row_index``: source row of each
        output row (``None`` = all rows in order); ``candidates``: ``(rows, k)`` dictionary rows, -1 = none."""
        import torch

        dev = torch.device("cuda", self.device)
        cand = torch.as_tensor(np.ascontiguousarray(candidates, dtype=np.int64)).to(dev)
        rows, k = cand.shape
        idx = None
        if exp_row_index is not None:
            idx = torch.as_tensor(np.ascontiguousarray(exp_row_index, dtype=np.int64)).to(dev)
        out = torch.empty((rows, k), dtype=torch.float64, device=dev)
        ecode, dcode = _TORCH_DTYPES[str(exp_rows_dev.dtype)], _TORCH_DTYPES[str(dict_rows_dev.dtype)]
        S = int(exp_rows_dev.shape[1])
        if int(dict_rows_dev.shape[1]) != S:
            raise ValueError(f"Experimental ({S}) and dictionary ({int(dict_rows_dev.shape[1])}) signal sizes must be identical")
        self._stream_sync(dev)
        self._check(self._lib.kdi_scores_f64(
            self._h, exp_rows_dev.data_ptr(), ecode, idx.data_ptr() if idx is not None else None, rows,
            dict_rows_dev.data_ptr(), dcode, int(dict_rows_dev.shape[0]), S, int(metric), cand.data_ptr(), int(k),
            out.data_ptr()))
        return out.cpu().numpy()

    def merge_topk(self, scores, indices, k_out: int):
        """Merge ``(n_lists, rows, k_in)`` ranked lists into ``(rows, k_out)``.

        NumPy arrays are staged through the device; CUDA torch tensors are merged in place on
        the device and CUDA tensors are returned.
        """
        if hasattr(scores, "is_cuda") and scores.is_cuda:
            import torch

            n_lists, rows, k_in = scores.shape
            scores = scores.contiguous()
            indices = indices.contiguous()
            so = torch.empty((rows, k_out), dtype=torch.float32, device=scores.device)
            io = torch.empty((rows, k_out), dtype=torch.int64, device=scores.device)
            self._stream_sync(scores.device)
            self._check(
                self._lib.kdi_merge_topk(
                    self._h, rows, n_lists, k_in, scores.data_ptr(), indices.data_ptr(), k_out,
                    so.data_ptr(), io.data_ptr(), KDI_DEVICE,
                )
            )
            return io, so
        s = np.ascontiguousarray(scores, dtype=np.float32)
        i = np.ascontiguousarray(indices, dtype=np.int64)
        n_lists, rows, k_in = s.shape
        so = np.empty((rows, k_out), dtype=np.float32)
        io = np.empty((rows, k_out), dtype=np.int64)
        self._check(
            self._lib.kdi_merge_topk(
                self._h, rows, n_lists, k_in, s.ctypes.data, i.ctypes.data, k_out, so.ctypes.data,
                io.ctypes.data, KDI_HOST,
            )
        )
        return io, so

    def dictionary_indexing(
        self, experimental, exp_rows: int, dictionary, dict_rows: int, metric: int, keep_n: int,
        n_per_iteration: int = 0, nav_mask=None, index_offset: int = 0, out=None,
    ):
        """The whole driver on raw buffers (host NumPy / pinned, or CUDA torch tensors)."""
        eptr, eloc, ecode, ekeep = _buffer(experimental, self)
        dptr, dloc, dcode, dkeep = _buffer(dictionary, self)
        n_e = int(np.prod(experimental.shape))
        n_d = int(np.prod(dictionary.shape))
        if exp_rows < 1 or n_e % exp_rows or dict_rows < 1 or n_d % dict_rows:
            raise ValueError("pattern arrays cannot be reshaped to (rows, -1)")
        S = n_e // exp_rows
        if n_d // dict_rows != S:
            raise ValueError(f"Experimental ({S}) and dictionary ({n_d // dict_rows}) signal sizes must be identical")
        rm = None
        kept = exp_rows
        if nav_mask is not None:
            rm = np.ascontiguousarray(np.asarray(nav_mask).ravel().astype(np.uint8))
            kept = int((rm == 0).sum())
        if out is None:
            scores = np.empty((kept, keep_n), dtype=np.float32)
            idx = np.empty((kept, keep_n), dtype=np.int64)
            sptr, iptr, oloc = scores.ctypes.data, idx.ctypes.data, KDI_HOST
        else:
            idx, scores = out
            if hasattr(scores, "is_cuda") and scores.is_cuda:
                sptr, iptr, oloc = scores.data_ptr(), idx.data_ptr(), KDI_DEVICE
            else:
                sptr, iptr, oloc = scores.ctypes.data, idx.ctypes.data, KDI_HOST
        self._check(
            self._lib.kdi_dictionary_indexing(
                self._h, eptr, eloc, ecode, exp_rows, dptr, dloc, dcode, dict_rows, S, metric, keep_n,
                int(n_per_iteration or 0), rm.ctypes.data if rm is not None else None, index_offset,
                sptr, iptr, oloc,
            )
        )
        del ekeep, dkeep
        return idx, scores

    def indexing_job(self, experimental, exp_rows: int, dict_rows: int, metric: int, keep_n: int,
                     nav_mask=None, index_offset: int = 0) -> "IndexingJob":
        """Start an appendable job (``kdi_job_begin``): the dictionary follows chunk by chunk through
        :meth:`IndexingJob.append` - the reference's loop over ``dictionary[start:end]``
        (``_dictionary_indexing.py:102-128``) with the device working behind the caller."""
        eptr, eloc, ecode, ekeep = _buffer(experimental, self)
        n_e = int(np.prod(experimental.shape))
        if exp_rows < 1 or n_e % exp_rows:
            raise ValueError("pattern array cannot be reshaped to (rows, -1)")
        S = n_e // exp_rows
        rm, kept = None, exp_rows
        if nav_mask is not None:
            rm = np.ascontiguousarray(np.asarray(nav_mask).ravel().astype(np.uint8))
            kept = int((rm == 0).sum())
        h = _vp()
        self._check(self._lib.kdi_job_begin(
            self._h, eptr, eloc, ecode, exp_rows, int(dict_rows), S, metric, int(keep_n),
            rm.ctypes.data if rm is not None else None, int(index_offset), C.byref(h)))
        del ekeep
        return IndexingJob(self, h.value, kept, int(keep_n), S, int(dict_rows))

    # -- sharded dictionary: candidate pipeline (CUDA torch tensors in and out) -------------------
    def candidate_capacity(self, keep_n: int) -> int:
        return int(self._lib.kdi_candidate_capacity_ctx(self._h, int(keep_n)))

    def certificate_bound(self, row_length: int, compute_dtype: int = 0) -> float:
        """The bound of the strict certificate (``OPT_CERT_STRICT``) on abs(tensor-core score - float32
        score) for rows of ``row_length`` kept values."""
        return float(self._lib.kdi_certificate_bound(int(compute_dtype), int(row_length)))

    def shard_candidates(self, experimental, exp_rows, dictionary, dict_rows, metric, keep_n,
                         nav_mask=None, index_offset=0, pad_rows=0):
        """Stage 1 on this rank's dictionary rows.  Returns ``(shard, approx, gidx)``: CUDA tensors
        ``(max(rows kept, pad_rows), kc)`` with the best candidates by tensor-core score and global
        indices; rows past the kept ones are padding (-inf / -1)."""
        import torch

        kc = self.candidate_capacity(keep_n)
        if kc == 0:
            raise NotImplementedError(f"keep_n {keep_n} too large for the candidate pipeline")
        eptr, eloc, ecode, ekeep = _buffer(experimental, self)
        dptr, dloc, dcode, dkeep = _buffer(dictionary, self)
        n_e = int(np.prod(experimental.shape))
        n_d = int(np.prod(dictionary.shape))
        if exp_rows < 1 or n_e % exp_rows or dict_rows < 1 or n_d % dict_rows:
            raise ValueError("pattern arrays cannot be reshaped to (rows, -1)")
        S = n_e // exp_rows
        if n_d // dict_rows != S:
            raise ValueError(f"Experimental ({S}) and dictionary ({n_d // dict_rows}) signal sizes must be identical")
        rm, kept = None, exp_rows
        if nav_mask is not None:
            rm = np.ascontiguousarray(np.asarray(nav_mask).ravel().astype(np.uint8))
            kept = int((rm == 0).sum())
        dev = torch.device("cuda", self.device)
        n_out = max(kept, int(pad_rows))
        approx = torch.empty((n_out, kc), dtype=torch.float32, device=dev)
        gidx = torch.empty((n_out, kc), dtype=torch.int64, device=dev)
        if n_out > kept:
            approx[kept:].fill_(-float("inf"))
            gidx[kept:].fill_(-1)
            self._stream_sync(dev)
        h = _vp()
        self._check(
            self._lib.kdi_shard_candidates(
                self._h, eptr, eloc, ecode, exp_rows, dptr, dloc, dcode, dict_rows, S, metric, int(keep_n),
                rm.ctypes.data if rm is not None else None, int(index_offset), approx.data_ptr(),
                gidx.data_ptr(), C.byref(h),
            )
        )
        del ekeep
        return Shard(self, h.value, kept, kc, keep=(dictionary, dkeep)), approx, gidx

    def comm_bytes_needed(self, world: int, rows: int, keep_n: int) -> int:
        kc = self.candidate_capacity(keep_n)
        if kc == 0:
            raise NotImplementedError(f"keep_n {keep_n} too large for the candidate pipeline")
        n = int(self._lib.kdi_comm_bytes_needed(int(world), int(rows), kc, int(keep_n)))
        if n < 0:
            raise ValueError("bad shape for the peer exchange")
        return n

    def shard_run_peer(self, comm: PeerComm, experimental, exp_rows, dictionary, dict_rows, metric, keep_n,
                       dict_total, nav_mask=None):
        """The whole sharded job with the exchange over peer-mapped memory (``kdi_shard_run_peer``;
        collective: every rank calls it with the same shapes).  ``dictionary``: this rank's rows (array
        / CUDA tensor), or a ``(master pattern, rotations)`` pair for a generated shard.  Returns
        ``(shard, indices, scores, flagged rows)`` - CUDA tensors, complete and identical on every rank."""
        import torch

        eptr, eloc, ecode, ekeep = _buffer(experimental, self)
        n_e = int(np.prod(experimental.shape))
        if exp_rows < 1 or n_e % exp_rows:
            raise ValueError("pattern arrays cannot be reshaped to (rows, -1)")
        S = n_e // exp_rows
        rm, kept = None, exp_rows
        if nav_mask is not None:
            rm = np.ascontiguousarray(np.asarray(nav_mask).ravel().astype(np.uint8))
            kept = int((rm == 0).sum())
        dev = torch.device("cuda", self.device)
        scores = torch.empty((kept, keep_n), dtype=torch.float32, device=dev)
        idx = torch.empty((kept, keep_n), dtype=torch.int64, device=dev)
        flags = torch.empty((max(kept, 1),), dtype=torch.int32, device=dev)
        n_flag = C.c_int(0)
        h = _vp()
        self._stream_sync(dev)
        rmp = rm.ctypes.data if rm is not None else None
        held = None
        if isinstance(dictionary, tuple):
            mp, rotations = dictionary
            rptr, rloc, rkeep, n = _rotations(rotations)
            if rloc == KDI_DEVICE:
                self._stream_sync(rkeep.device)
            self._check(self._lib.kdi_shard_run_peer_projected(
                self._h, comm._h, eptr, eloc, ecode, exp_rows, S, mp._h, rptr, rloc, n, metric, int(keep_n), rmp,
                int(dict_total), scores.data_ptr(), idx.data_ptr(), flags.data_ptr(), C.byref(n_flag), C.byref(h)))
            del rkeep
        else:
            dptr, dloc, dcode, dkeep = _buffer(dictionary, self)
            n_d = int(np.prod(dictionary.shape))
            if dict_rows < 1 or n_d % dict_rows or n_d // dict_rows != S:
                raise ValueError(f"Experimental ({S}) and dictionary signal sizes must be identical")
            self._check(self._lib.kdi_shard_run_peer(
                self._h, comm._h, eptr, eloc, ecode, exp_rows, dptr, dloc, dcode, dict_rows, S, metric, int(keep_n), rmp,
                int(dict_total), scores.data_ptr(), idx.data_ptr(), flags.data_ptr(), C.byref(n_flag), C.byref(h)))
            held = (dictionary, dkeep)
        del ekeep
        kc = self.candidate_capacity(keep_n)
        return Shard(self, h.value, kept, kc, keep=held), idx, scores, flags[: n_flag.value]

    # -- dictionary generation -------------------------------------------------------------------
    def master_pattern(self, upper, lower, direction_cosines, scale=None, rescale=False, out_min=-1.0,
                       out_max=1.0) -> MasterPattern:
        up = np.ascontiguousarray(upper)
        lo = np.ascontiguousarray(lower)
        if up.dtype not in _DTYPES:
            up, lo = up.astype(np.float64), lo.astype(np.float64)
        if lo.dtype != up.dtype or lo.shape != up.shape or up.ndim != 2:
            raise ValueError("master pattern hemispheres must be 2-D arrays of the same shape and dtype")
        dc = np.ascontiguousarray(np.asarray(direction_cosines, dtype=np.float64).reshape(-1, 3))
        if scale is None:
            scale = (up.shape[1] - 1) / 2  # ebsd_master_pattern.py:255-256
        h = _vp()
        self._check(
            self._lib.kdi_master_pattern_create(
                self._h, up.ctypes.data, lo.ctypes.data, _DTYPES[up.dtype], up.shape[0], up.shape[1],
                dc.ctypes.data, dc.shape[0], float(scale), int(bool(rescale)), float(out_min), float(out_max),
                C.byref(h),
            )
        )
        return MasterPattern(self, h.value, dc.shape[0])

    def project_patterns(self, mp: MasterPattern, rotations, out=None):
        """``(n, S)`` float32 patterns (NumPy, or the CUDA tensor ``out``)."""
        rptr, rloc, keep, n = _rotations(rotations)
        if out is None:
            res = np.empty((n, mp.n_pixels), dtype=np.float32)
            optr, oloc = res.ctypes.data, KDI_HOST
        else:
            res = out
            self._stream_sync(out.device)
            optr, oloc = out.data_ptr(), KDI_DEVICE
        if rloc == KDI_DEVICE:
            self._stream_sync(keep.device)
        self._check(self._lib.kdi_project_patterns(self._h, mp._h, rptr, rloc, n, optr, oloc))
        del keep
        return res

    def project_patterns_varying_pc(self, mp: MasterPattern, rotations, pcs, nrows, ncols, om_detector_to_sample):
        """``(n, nrows * ncols)`` float32 patterns, rotation ``i`` seen from projection centre ``pcs[i]``."""
        rot = np.ascontiguousarray(rotations, dtype=np.float64).reshape(-1, 4)
        pc = np.ascontiguousarray(pcs, dtype=np.float64).reshape(-1, 3)
        if pc.shape[0] != rot.shape[0]:
            raise ValueError("one projection centre per rotation is needed")
        om = np.ascontiguousarray(om_detector_to_sample, dtype=np.float64).reshape(3, 3)
        res = np.empty((rot.shape[0], mp.n_pixels), dtype=np.float32)
        self._check(self._lib.kdi_project_patterns_varying_pc(
            self._h, mp._h, rot.ctypes.data, rot.shape[0], pc.ctypes.data, int(nrows), int(ncols), om.ctypes.data,
            res.ctypes.data))
        return res

    def patterns_projected(self, mp: MasterPattern, rotations, metric: int) -> Patterns:
        rptr, rloc, keep, n = _rotations(rotations)
        if rloc == KDI_DEVICE:
            self._stream_sync(keep.device)
        h = _vp()
        self._check(self._lib.kdi_patterns_create_projected(self._h, mp._h, rptr, rloc, n, metric, C.byref(h)))
        del keep
        return Patterns(self, h.value)

    def dictionary_indexing_projected(self, experimental, exp_rows: int, mp: MasterPattern, rotations, metric: int,
                                      keep_n: int, nav_mask=None, index_offset: int = 0, out=None):
        """The whole driver with the dictionary generated on the device from ``rotations``."""
        eptr, eloc, ecode, ekeep = _buffer(experimental, self)
        rptr, rloc, rkeep, n = _rotations(rotations)
        if rloc == KDI_DEVICE:
            self._stream_sync(rkeep.device)
        n_e = int(np.prod(experimental.shape))
        if exp_rows < 1 or n_e % exp_rows:
            raise ValueError("pattern array cannot be reshaped to (rows, -1)")
        S = n_e // exp_rows
        if S != mp.n_pixels:
            raise ValueError(f"Experimental ({S}) and dictionary ({mp.n_pixels}) signal sizes must be identical")
        rm, kept = None, exp_rows
        if nav_mask is not None:
            rm = np.ascontiguousarray(np.asarray(nav_mask).ravel().astype(np.uint8))
            kept = int((rm == 0).sum())
        if out is None:
            scores = np.empty((kept, keep_n), dtype=np.float32)
            idx = np.empty((kept, keep_n), dtype=np.int64)
            sptr, iptr, oloc = scores.ctypes.data, idx.ctypes.data, KDI_HOST
        else:
            idx, scores = out
            if hasattr(scores, "is_cuda") and scores.is_cuda:
                sptr, iptr, oloc = scores.data_ptr(), idx.data_ptr(), KDI_DEVICE
            else:
                sptr, iptr, oloc = scores.ctypes.data, idx.ctypes.data, KDI_HOST
        self._check(
            self._lib.kdi_dictionary_indexing_projected(
                self._h, eptr, eloc, ecode, exp_rows, S, mp._h, rptr, rloc, n, metric, keep_n,
                rm.ctypes.data if rm is not None else None, index_offset, sptr, iptr, oloc,
            )
        )
        del ekeep, rkeep
        return idx, scores

    def shard_candidates_projected(self, experimental, exp_rows, mp: MasterPattern, rotations, metric, keep_n,
                                   nav_mask=None, index_offset=0, pad_rows=0):
        """``shard_candidates`` with this rank's dictionary rows generated from its rotations."""
        import torch

        kc = self.candidate_capacity(keep_n)
        if kc == 0:
            raise NotImplementedError(f"keep_n {keep_n} too large for the candidate pipeline")
        eptr, eloc, ecode, ekeep = _buffer(experimental, self)
        rptr, rloc, rkeep, n = _rotations(rotations)
        if rloc == KDI_DEVICE:
            self._stream_sync(rkeep.device)
        n_e = int(np.prod(experimental.shape))
        if exp_rows < 1 or n_e % exp_rows:
            raise ValueError("pattern array cannot be reshaped to (rows, -1)")
        S = n_e // exp_rows
        rm, kept = None, exp_rows
        if nav_mask is not None:
            rm = np.ascontiguousarray(np.asarray(nav_mask).ravel().astype(np.uint8))
            kept = int((rm == 0).sum())
        dev = torch.device("cuda", self.device)
        n_out = max(kept, int(pad_rows))
        approx = torch.empty((n_out, kc), dtype=torch.float32, device=dev)
        gidx = torch.empty((n_out, kc), dtype=torch.int64, device=dev)
        if n_out > kept:
            approx[kept:].fill_(-float("inf"))
            gidx[kept:].fill_(-1)
            self._stream_sync(dev)
        h = _vp()
        self._check(
            self._lib.kdi_shard_candidates_projected(
                self._h, eptr, eloc, ecode, exp_rows, S, mp._h, rptr, rloc, n, metric, int(keep_n),
                rm.ctypes.data if rm is not None else None, int(index_offset), approx.data_ptr(),
                gidx.data_ptr(), C.byref(h),
            )
        )
        del ekeep, rkeep
        return Shard(self, h.value, kept, kc), approx, gidx

    def orientation_similarity_map(self, indices, ny, nx, n_best, from_n_best, normalize, footprint, center_index):
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        keep_n = idx.shape[1]
        fp = np.ascontiguousarray(np.asarray(footprint).astype(bool).astype(np.uint8))
        out = np.empty((ny, nx, n_best - from_n_best + 1), dtype=np.float32)
        self._check(
            self._lib.kdi_orientation_similarity_map(
                self._h, idx.ctypes.data, ny, nx, keep_n, n_best, from_n_best, int(bool(normalize)),
                fp.ctypes.data, fp.shape[0], fp.shape[1], center_index, out.ctypes.data,
            )
        )
        return out

    def _pattern_io(self, patterns, device_output):
        """(pointer, location, dtype code, keep-alive, output array/tensor, its pointer, its location)."""
        ptr, loc, code, keep = _buffer(patterns, self)
        if code not in (KDI_U8, KDI_U16, KDI_F32):
            raise NotImplementedError("patterns must be uint8, uint16 or float32")
        if loc == KDI_DEVICE or device_output:
            import torch

            dev = keep.device if loc == KDI_DEVICE else torch.device("cuda", self.device)
            tdt = {KDI_U8: torch.uint8, KDI_U16: torch.uint16, KDI_F32: torch.float32}[code]
            out = torch.empty(tuple(keep.shape), dtype=tdt, device=dev)
            self._stream_sync(dev)
            return ptr, loc, code, keep, out, out.data_ptr(), KDI_DEVICE
        out = np.empty(keep.shape, dtype=keep.dtype)
        return ptr, loc, code, keep, out, out.ctypes.data, KDI_HOST

    def preprocess_patterns(self, patterns, nrows, ncols, static_op=0, static_bg=None, scale_bg=False, dynamic_op=0,
                            dynamic_domain=0, weights_y=None, weights_x=None, device_output=False):
        """``kdi_preprocess_patterns`` on ``(n, nrows * ncols)`` patterns (NumPy or CUDA tensor)."""
        ptr, loc, code, keep, out, optr, oloc = self._pattern_io(patterns, device_output)
        n = int(np.prod(keep.shape)) // (nrows * ncols)
        bg = None if static_bg is None else np.ascontiguousarray(static_bg, dtype=np.float32).reshape(-1)
        wy = None if weights_y is None else np.ascontiguousarray(weights_y, dtype=np.float64)
        wx = None if weights_x is None else np.ascontiguousarray(weights_x, dtype=np.float64)
        self._check(
            self._lib.kdi_preprocess_patterns(
                self._h, ptr, loc, code, n, int(nrows), int(ncols), int(static_op),
                None if bg is None else bg.ctypes.data, int(bool(scale_bg)), int(dynamic_op), int(dynamic_domain),
                None if wy is None else wy.ctypes.data, 0 if wy is None else wy.size,
                None if wx is None else wx.ctypes.data, 0 if wx is None else wx.size, optr, oloc,
            )
        )
        del keep
        return out

    def average_neighbour_patterns(self, patterns, ny, nx, n_pixels, window, window_sums, device_output=False):
        """``kdi_average_neighbour_patterns`` on the ``(ny * nx, n_pixels)`` patterns of a map."""
        ptr, loc, code, keep, out, optr, oloc = self._pattern_io(patterns, device_output)
        w = np.ascontiguousarray(window, dtype=np.float64)
        sums = np.ascontiguousarray(window_sums, dtype=np.int32)
        self._check(
            self._lib.kdi_average_neighbour_patterns(
                self._h, ptr, loc, code, int(ny), int(nx), int(n_pixels), w.ctypes.data, w.shape[0], w.shape[1],
                sums.ctypes.data, optr, oloc,
            )
        )
        del keep
        return out

    def refine(self, mp: "MasterPattern", mode: int, patterns, nrows: int, ncols: int, rescale: bool, x0,
               lower=None, upper=None, rotations=None, pcs=None, om_detector_to_sample=None, xatol=1e-4,
               fatol=1e-4, maxiter=-1, maxfev=-1, adaptive=False) -> np.ndarray:
        """``kdi_refine``: Nelder-Mead refinement of ``x0`` ``(n, starts, 3 | 6)`` for ``patterns``
        ``(n, nrows * ncols)`` (NumPy array or CUDA tensor).  Returns the reference's result rows
        ``(n, 2 + n_var [+ 1])``: score, evaluations, variables[, best start]."""
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        n, n_starts, nv = x0.shape
        ptr, loc, code, keep = _buffer(patterns, self)

        def opt(a, shape):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != shape:
                raise ValueError(f"expected an array of shape {shape}, got {a.shape}")
            return a

        lo, hi = opt(lower, x0.shape), opt(upper, x0.shape)
        rot, pc = opt(rotations, (n, 4)), opt(pcs, (n, 3))
        om = opt(om_detector_to_sample, (3, 3))
        o = RefineOptions(float(xatol), float(fatol), int(maxiter), int(maxfev), int(bool(adaptive)))
        out = np.empty((n, 2 + nv + (1 if n_starts > 1 else 0)), dtype=np.float64)

        def p(a):
            return None if a is None else a.ctypes.data

        self._check(
            self._lib.kdi_refine(
                self._h, mp._h, int(mode), ptr, loc, code, n, int(nrows), int(ncols), int(bool(rescale)),
                x0.ctypes.data, n_starts, p(lo), p(hi), p(rot), p(pc), p(om), C.byref(o), out.ctypes.data,
            )
        )
        del keep
        return out

    def refine_objective(self, mp: "MasterPattern", mode: int, patterns, nrows: int, ncols: int, rescale: bool,
                         pattern_rows, x, rotations=None, pcs=None, om_detector_to_sample=None) -> np.ndarray:
        """``kdi_refine_objective``: ``1 - NCC`` for row ``i`` = pattern ``pattern_rows[i]`` at each of the
        parameter sets ``x[i]`` ``(rows, points, 3 | 6)``.  ``patterns``: ``(n, nrows * ncols)`` NumPy array
        or (kept on the device between calls) CUDA tensor.  Returns ``(rows, points)`` float64."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        rows, n_points, nv = x.shape
        ptr, loc, code, keep = _buffer(patterns, self)
        n_src = int(patterns.shape[0])
        pr = np.ascontiguousarray(pattern_rows, dtype=np.int64)
        if pr.shape != (rows,):
            raise ValueError(f"expected {rows} pattern rows, got {pr.shape}")
        rot = None if rotations is None else np.ascontiguousarray(rotations, dtype=np.float64).reshape(rows, n_points, 4)
        pc = None if pcs is None else np.ascontiguousarray(pcs, dtype=np.float64).reshape(rows, 3)
        om = None if om_detector_to_sample is None else np.ascontiguousarray(om_detector_to_sample, dtype=np.float64).reshape(3, 3)
        out = np.empty((rows, n_points), dtype=np.float64)

        def p(a):
            return None if a is None else a.ctypes.data

        self._check(self._lib.kdi_refine_objective(
            self._h, mp._h, int(mode), ptr, loc, code, n_src, int(nrows), int(ncols), int(bool(rescale)),
            pr.ctypes.data, rows, x.ctypes.data, n_points, p(rot), p(pc), p(om), out.ctypes.data))
        del keep
        return out

    def merge_crystal_maps(self, scores, rotations, simulation_indices, point_rows, not_indexed,
                           map_size, mean_n_best, sign, idx_as_double):
        """``kdi_merge_crystal_maps``: per-map lists of ``(n_i, N)`` scores (float32 or float64),
        ``(n_i, N, 4)`` rotations, optional ``(n_i, N)`` simulation indices, optional int32 row
        maps and not-indexed flags.  Returns a dict of the merged arrays."""
        n_maps = len(scores)
        dt = np.dtype(scores[0].dtype)
        if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError(f"scores must be float32 or float64, not {dt}")
        sc = [np.ascontiguousarray(s, dtype=dt) for s in scores]
        n_scores = sc[0].shape[1]
        rot = [np.ascontiguousarray(r, dtype=np.float64) for r in rotations]
        idx = None if simulation_indices is None else [np.ascontiguousarray(i, dtype=np.int64) for i in simulation_indices]
        rows = [None if r is None else np.ascontiguousarray(r, dtype=np.int32) for r in point_rows]
        ni = [None if f is None else np.ascontiguousarray(f, dtype=np.uint8) for f in not_indexed]

        def ptrs(arrs):
            return (C.c_void_p * n_maps)(*[None if a is None else a.ctypes.data for a in arrs])

        n_pts = (C.c_int64 * n_maps)(*[a.shape[0] for a in sc])
        total = n_scores * n_maps
        out = {
            "phase_id": np.empty(map_size, dtype=np.int64),
            "scores": np.empty((map_size, n_scores), dtype=dt),
            "rotations": np.empty((map_size, n_scores, 4), dtype=np.float64),
            "merged_scores": np.empty((map_size, total), dtype=dt),
        }
        if idx is not None:
            out["simulation_indices"] = np.empty((map_size, n_scores), dtype=np.int32)
            out["merged_simulation_indices"] = np.empty((map_size, total), dtype=np.float64 if idx_as_double else np.int64)
        self._check(
            self._lib.kdi_merge_crystal_maps(
                self._h, n_maps, int(map_size), int(n_scores), _DTYPES[dt], n_pts, ptrs(sc), ptrs(rot),
                None if idx is None else ptrs(idx),
                ptrs(rows) if any(r is not None for r in rows) else None,
                ptrs(ni) if any(f is not None for f in ni) else None,
                int(mean_n_best), int(sign), int(bool(idx_as_double)), out["phase_id"].ctypes.data,
                out["scores"].ctypes.data, out["rotations"].ctypes.data,
                out["simulation_indices"].ctypes.data if idx is not None else None,
                out["merged_scores"].ctypes.data,
                out["merged_simulation_indices"].ctypes.data if idx is not None else None,
            )
        )
        return out


def bind_to_gpu_numa_node(device: int) -> list[int] | None:
    """Pin this process to the CPUs that are local to ``device`` (NVML's CPU affinity mask), so
    that pinned host buffers allocated afterwards are first-touched on the GPU's own NUMA node.
    With one process per GPU this keeps each rank's H2D traffic off the inter-socket link (on an
    8-GPU box the aggregate upload rate is otherwise bounded by one socket's memory).  Returns the
    CPU list, or ``None`` when NVML or the affinity call is unavailable (nothing is changed)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        n_cpus = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpus + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001 - best effort: NVML missing, container without the call, ...
        return None


_default_ctx: dict[int, Context] = {}


def default_context(device: int | None = None) -> Context:
    """Process-wide context for ``device`` (default: ``LOCAL_RANK`` or 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    ctx = _default_ctx.get(device)
    if ctx is None or ctx._h is None:
        ctx = Context(device)
        _default_ctx[device] = ctx
    return ctx
