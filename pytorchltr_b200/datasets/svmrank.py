"""SVMrank text ingestion on ``csrc/libltr_svmrank.so``, a multithreaded C++ parser.

``parse_svmrank_file(path)`` mirrors the reference's Cython entry point
(pytorchltr/datasets/svmrank/parser/svmrank_parser.pyx:19-59 over svmrank_parser.h:174-515): same grammar,
same ``(xs float64 (rows, cols), ys int32 (rows,), qids int64 (rows,))`` result, same exceptions
(``OSError`` for an unreadable file, ``ValueError`` for a file that is not in SVMrank format).
``load_svmrank(path)`` goes straight to a :class:`DeviceRankingDataset` (float32 features, query offsets from
runs of equal qids as in svmrank.py:69-72), dense or CSR.
"""
import ctypes
import os
import threading
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "csrc", "libltr_svmrank.so")
SYMBOLS = ("ltr_svmrank_parse", "ltr_svmrank_rows", "ltr_svmrank_cols", "ltr_svmrank_nnz", "ltr_svmrank_fill_f64",
           "ltr_svmrank_fill_f32", "ltr_svmrank_fill_csr", "ltr_svmrank_release")
PARSE_OK, PARSE_FILE_ERROR, PARSE_FORMAT_ERROR, PARSE_MEMORY_ERROR = 0, 1, 2, 3

_lock = threading.Lock()
_lib = None


def lib():
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(f"{LIB_PATH} is missing: run `python -m pytorchltr_b200.build`")
                h = ctypes.CDLL(LIB_PATH, use_errno=True)
                vp, ci, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64
                h.ltr_svmrank_parse.restype = ci
                h.ltr_svmrank_parse.argtypes = [ctypes.c_char_p, ci, ctypes.POINTER(vp)]
                for name in ("ltr_svmrank_rows", "ltr_svmrank_cols", "ltr_svmrank_nnz"):
                    getattr(h, name).restype = u64
                    getattr(h, name).argtypes = [vp]
                for name in ("ltr_svmrank_fill_f64", "ltr_svmrank_fill_f32"):
                    getattr(h, name).restype = ci
                    getattr(h, name).argtypes = [vp, vp, vp, vp, ci]
                h.ltr_svmrank_fill_csr.restype = ci
                h.ltr_svmrank_fill_csr.argtypes = [vp, vp, vp, vp]
                h.ltr_svmrank_release.restype = None
                h.ltr_svmrank_release.argtypes = [vp]
                _lib = h
    return _lib


class _Parsed:
    """A parsed file held by the library until its arrays have been filled."""

    def __init__(self, path: str, n_threads: int):
        self.h = lib()
        self.handle = ctypes.c_void_p()
        rc = self.h.ltr_svmrank_parse(os.fsencode(path), int(n_threads), ctypes.byref(self.handle))
        if rc == PARSE_FILE_ERROR:
            raise OSError(ctypes.get_errno(), "could not open file %s" % path)
        if rc == PARSE_FORMAT_ERROR:
            raise ValueError("could not parse file %s, not in SVMrank format" % path)
        if rc != PARSE_OK:
            raise OSError(ctypes.get_errno(), "could not allocate memory")
        self.rows = int(self.h.ltr_svmrank_rows(self.handle))
        self.cols = int(self.h.ltr_svmrank_cols(self.handle))
        self.nnz = int(self.h.ltr_svmrank_nnz(self.handle))
        if self.rows == 0:
            # the reference cannot wrap an empty result either (ValueError from its array views)
            self.close()
            raise ValueError("could not parse file %s, not in SVMrank format" % path)

    def close(self):
        if self.handle:
            self.h.ltr_svmrank_release(self.handle)
            self.handle = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def parse_svmrank_file(path: str, dtype=np.float64, n_threads: int = 0):
    """``(xs, ys, qids)`` of an SVMrank file; ``dtype`` float64 (the reference's) or float32;
    ``n_threads`` 0 = one per hardware thread."""
    dtype = np.dtype(dtype)
    if dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
        raise ValueError("dtype must be float64 or float32")
    with _Parsed(path, n_threads) as p:
        xs = np.empty((p.rows, p.cols), dtype=dtype)
        ys = np.empty(p.rows, dtype=np.int32)
        qids = np.empty(p.rows, dtype=np.int64)
        fill = p.h.ltr_svmrank_fill_f64 if dtype == np.float64 else p.h.ltr_svmrank_fill_f32
        rc = fill(p.handle, xs.ctypes.data, ys.ctypes.data, qids.ctypes.data, int(n_threads))
        if rc != PARSE_OK:
            raise RuntimeError("ltr_svmrank_fill failed")
    return xs, ys, qids


def query_offsets(qids: np.ndarray):
    """Offsets and unique qids of runs of equal qids (svmrank.py:69-72)."""
    if len(qids) == 0:
        return np.zeros(1, dtype=np.int64), qids[:0]
    offsets = np.hstack([[0], np.where(qids[1:] != qids[:-1])[0] + 1, [len(qids)]]).astype(np.int64)
    return offsets, qids[offsets[:-1]]


def load_svmrank(path: str, device=None, sparse: bool = False, filter_queries: bool = False,
                 n_threads: int = 0):
    """Parses ``path`` and returns a :class:`pytorchltr_b200.datasets.DeviceRankingDataset` on ``device``
    (float32 features; ``sparse=True`` keeps them in CSR form).  ``filter_queries`` drops the queries
    without a relevant document, like ``SVMRankDataset(filter_queries=True)`` (svmrank.py:87-96)."""
    import torch

    from pytorchltr_b200.datasets.device import DeviceRankingDataset
    with _Parsed(path, n_threads) as p:
        ys = np.empty(p.rows, dtype=np.int32)
        qids = np.empty(p.rows, dtype=np.int64)
        if sparse:
            indptr = np.empty(p.rows + 1, dtype=np.int64)
            indices = np.empty(p.nnz, dtype=np.int64)
            values = np.empty(p.nnz, dtype=np.float32)
            p.h.ltr_svmrank_fill_f32(p.handle, None, ys.ctypes.data, qids.ctypes.data, int(n_threads))
            p.h.ltr_svmrank_fill_csr(p.handle, indptr.ctypes.data, indices.ctypes.data, values.ctypes.data)
        else:
            xs = np.empty((p.rows, p.cols), dtype=np.float32)
            p.h.ltr_svmrank_fill_f32(p.handle, xs.ctypes.data, ys.ctypes.data, qids.ctypes.data, int(n_threads))
        cols = p.cols
    offsets, unique = query_offsets(qids)
    if filter_queries and len(unique):
        rel_sum = np.add.reduceat(ys.astype(np.int64), offsets[:-1])
        keep = rel_sum > 0
        if not keep.all():
            doc_keep = np.repeat(keep, np.diff(offsets))
            counts = np.diff(offsets)[keep]
            if sparse:
                row_nnz = np.diff(indptr)
                nz_keep = np.repeat(doc_keep, row_nnz)
                indices, values = indices[nz_keep], values[nz_keep]
                indptr = np.concatenate([[0], np.cumsum(row_nnz[doc_keep])]).astype(np.int64)
            else:
                xs = xs[doc_keep]
            ys, unique = ys[doc_keep], unique[keep]
            offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    rel = torch.from_numpy(ys.astype(np.int64))
    off_t, qid_t = torch.from_numpy(offsets), torch.from_numpy(np.ascontiguousarray(unique))
    if sparse:
        return DeviceRankingDataset.from_csr(torch.from_numpy(indptr), torch.from_numpy(indices),
                                             torch.from_numpy(values), cols, rel, off_t, qid_t, device=device)
    return DeviceRankingDataset(torch.from_numpy(xs), rel, off_t, qid_t, device=device)
