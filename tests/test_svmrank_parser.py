"""SVMrank text parser (csrc/svmrank_parser.cpp) against what the UNMODIFIED reference
parser returned for the same files (tests/golden/svmrank/cases.npz, made by
tests/golden/make_svmrank_golden.py from pytorchltr/datasets/svmrank/parser/svmrank_parser.{h,pyx}):
values bit-identical (same decimal arithmetic), same shapes, same exceptions."""
import os

import numpy as np
import pytest

from pytorchltr_b200.datasets.svmrank import parse_svmrank_file, query_offsets

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "svmrank", "cases.npz"))
NAMES = sorted({k.split("__")[0] for k in GOLDEN.files})


def _write(tmp_path, name):
    path = tmp_path / (name + ".txt")
    path.write_bytes(GOLDEN[name + "__text"].tobytes())
    return str(path)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("threads", (1, 4))
def test_parser_matches_reference_parser(tmp_path, name, threads):
    path = _write(tmp_path, name)
    if int(GOLDEN[name + "__error"]):
        with pytest.raises(ValueError):
            parse_svmrank_file(path, n_threads=threads)
        return
    xs, ys, qids = parse_svmrank_file(path, n_threads=threads)
    assert xs.dtype == np.float64 and ys.dtype == np.int32 and qids.dtype == np.int64
    assert np.array_equal(xs, GOLDEN[name + "__xs"])          # bit-identical doubles
    assert np.array_equal(ys, GOLDEN[name + "__ys"])
    assert np.array_equal(qids, GOLDEN[name + "__qids"])
    x32, _, _ = parse_svmrank_file(path, dtype=np.float32, n_threads=threads)
    assert np.array_equal(x32, GOLDEN[name + "__xs"].astype(np.float32))


def test_missing_file_raises_oserror(tmp_path):
    with pytest.raises(OSError):
        parse_svmrank_file(str(tmp_path / "nope.txt"))


def test_large_file_is_thread_count_invariant(tmp_path):
    """~6 MB, cut into many slices: every thread count gives the same arrays; values agree with the
    numbers that were printed."""
    rng = np.random.default_rng(5)
    rows, F = 20000, 40
    X = np.round(rng.standard_normal((rows, F)), 5)
    y = rng.integers(0, 5, size=rows)
    q = np.sort(rng.integers(1, 900, size=rows))
    with open(tmp_path / "big.txt", "w") as f:
        for i in range(rows):
            f.write("%d qid:%d %s\n" % (y[i], q[i], " ".join("%d:%.5f" % (c + 1, X[i, c]) for c in range(F))))
    ref = parse_svmrank_file(str(tmp_path / "big.txt"), n_threads=1)
    assert ref[0].shape == (rows, F)
    assert np.allclose(ref[0], X, rtol=0, atol=1e-12)
    assert np.array_equal(ref[1], y) and np.array_equal(ref[2], q)
    for threads in (2, 3, 8, 0):
        got = parse_svmrank_file(str(tmp_path / "big.txt"), n_threads=threads)
        assert all(np.array_equal(a, b) for a, b in zip(ref, got))
    offsets, unique = query_offsets(ref[2])
    assert offsets[0] == 0 and offsets[-1] == rows and np.array_equal(unique, np.unique(q))


@pytest.mark.gpu
def test_load_svmrank_to_device_dataset(tmp_path):
    """Text file -> DeviceRankingDataset (dense and CSR) -> collated batch, against the parsed arrays."""
    import torch
    from pytorchltr_b200.datasets.svmrank import load_svmrank
    path = _write(tmp_path, "sparse_mixed")
    xs, ys, qids = GOLDEN["sparse_mixed__xs"], GOLDEN["sparse_mixed__ys"], GOLDEN["sparse_mixed__qids"]
    offsets, unique = query_offsets(qids)
    dense = load_svmrank(path)
    sp = load_svmrank(path, sparse=True)
    assert len(dense) == len(unique) == len(sp)
    idx = list(range(len(unique)))
    a, b = dense.collate(idx), sp.collate(idx)
    L = int(np.diff(offsets).max())
    want = np.zeros((len(unique), L, xs.shape[1]), dtype=np.float32)
    rel = np.zeros((len(unique), L), dtype=np.int64)
    for i in range(len(unique)):
        n = offsets[i + 1] - offsets[i]
        want[i, :n] = xs[offsets[i]:offsets[i + 1]].astype(np.float32)
        rel[i, :n] = ys[offsets[i]:offsets[i + 1]]
    assert np.array_equal(a.features.cpu().numpy(), want)
    assert np.array_equal(b.features.to_dense().cpu().numpy(), want)
    assert np.array_equal(a.relevance.cpu().numpy(), rel) and torch.equal(a.relevance, b.relevance)
    assert np.array_equal(a.qid.cpu().numpy(), unique)
    filt = load_svmrank(path, filter_queries=True)
    keep = [i for i in range(len(unique)) if ys[offsets[i]:offsets[i + 1]].sum() > 0]
    assert len(filt) == len(keep) and np.array_equal(filt.qids.numpy(), unique[keep])


def _both_tiers(path, threads):
    """(fast-tier result or exception type, slow-tier result or exception type)"""
    out = []
    for slow in ("0", "1"):
        os.environ["LTR_SVMRANK_SLOW"] = slow
        try:
            out.append(parse_svmrank_file(path, n_threads=threads))
        except Exception as e:  # noqa: BLE001
            out.append(type(e))
        finally:
            os.environ.pop("LTR_SVMRANK_SLOW", None)
    return out


def _same(a, b):
    if isinstance(a, type) or isinstance(b, type):
        return a is b
    return all(x.shape == y.shape and x.dtype == y.dtype and np.array_equal(x.view(np.uint8), y.view(np.uint8))
               for x, y in zip(a, b))


HOSTILE_TOKENS = ["1:0", "2:7", "3:-4.25", "4:--3.5", "5:12345678", "6:123456789", "7:0.12345678", "8:0.123456789",
                  "9:1.5e3", "10:2.5E-2", "11:3.25e+1", "12:00012.5000", "13:99999999.99999999", "14:-0", "15:-0.0",
                  "99999:1", "16:1234567.12345678", "18:4.5e-", "19:1.0e+"]   # (column ids stay small: dense output)
BROKEN_TOKENS = ["17:3.e", "20:", "21:.", "22:1.", "23:a", "24:1x", "25:1e5", ":5", "26 :5", "27:-", "28:1..2", "29:1.2.3", "x:1"]


def _line(rng, tokens, y=None, qid=None):
    y = rng.integers(0, 5) if y is None else y
    qid = rng.integers(1, 50) if qid is None else qid
    sep = lambda: " " * int(rng.choice([1, 1, 1, 2, 3]))  # noqa: E731
    return f"{y}{sep()}qid:{qid}" + "".join(sep() + t for t in tokens)


def test_fast_and_slow_tiers_agree_on_hostile_files(tmp_path):
    """The two scanner tiers (SSE2 / SWAR token path, byte-at-a-time path) give the same bits -- or the same error --
    on files mixing every token form the grammar has, with comments, CRLF, blank runs, a missing final newline and,
    in a second family of files, exactly one malformed token at a random place."""
    rng = np.random.default_rng(7)
    for trial in range(60):
        lines = []
        for _ in range(int(rng.integers(1, 40))):
            k = int(rng.integers(0, 12))
            toks = [str(t) for t in rng.choice(HOSTILE_TOKENS, size=k)]
            toks += [f"{int(rng.integers(1, 300))}:{rng.random() * 10 ** int(rng.integers(-3, 6)):.{int(rng.integers(1, 9))}g}"
                     for _ in range(int(rng.integers(0, 20)))]
            toks = [t for t in toks if "e" not in t.split(":")[1] or "." in t.split(":")[1]]   # grammar: exponent needs a fraction
            rng.shuffle(toks)
            line = _line(rng, toks)
            r = rng.random()
            if r < 0.1:
                line += " # a comment with 5:6 in it"
            elif r < 0.2:
                line += "\r"
            elif r < 0.25:
                line = "# whole-line comment 1:2"
            lines.append(line)
        if trial % 2 == 1:                                  # one malformed token somewhere
            i = int(rng.integers(0, len(lines)))
            if not lines[i].startswith("#"):
                lines[i] = lines[i].split(" #")[0].rstrip("\r") + " " + str(rng.choice(BROKEN_TOKENS))
        text = "\n".join(lines) + ("\n" if rng.random() < 0.7 else "")
        path = tmp_path / f"hostile{trial}.txt"
        path.write_text(text)
        for threads in (1, 3):
            fast, slow = _both_tiers(str(path), threads)
            assert _same(fast, slow), (trial, threads, text[:400])


def test_csr_export_matches_the_dense_matrix(tmp_path):
    """ltr_svmrank_fill_csr (the sparse=True datasets) against the dense fill of the same parse: every stored value at
    its row / column, rows without features, several slices."""
    from pytorchltr_b200.datasets.svmrank import _Parsed
    rng = np.random.default_rng(3)
    lines = []
    for i in range(5000):
        cols = np.sort(rng.choice(np.arange(3, 60), size=int(rng.integers(0, 12)), replace=False))
        lines.append(f"{int(rng.integers(0, 5))} qid:{i // 25 + 1} " + " ".join(f"{c}:{rng.random() + 0.01:.5g}" for c in cols) + " ")
    path = tmp_path / "sparse.txt"
    path.write_text("\n".join(lines) + "\n")
    for threads in (1, 5):
        with _Parsed(str(path), threads) as p:
            xs = np.empty((p.rows, p.cols), dtype=np.float32)
            ys = np.empty(p.rows, dtype=np.int32)
            qids = np.empty(p.rows, dtype=np.int64)
            assert p.h.ltr_svmrank_fill_f32(p.handle, xs.ctypes.data, ys.ctypes.data, qids.ctypes.data, threads) == 0
            indptr = np.empty(p.rows + 1, dtype=np.int64)
            indices = np.empty(p.nnz, dtype=np.int64)
            values = np.empty(p.nnz, dtype=np.float32)
            assert p.h.ltr_svmrank_fill_csr(p.handle, indptr.ctypes.data, indices.ctypes.data, values.ctypes.data) == 0
        assert p.rows == 5000 and indptr[0] == 0 and indptr[-1] == len(values) == int((xs != 0).sum())
        dense = np.zeros_like(xs)
        rows = np.repeat(np.arange(5000), np.diff(indptr))
        dense[rows, indices] = values
        assert np.array_equal(dense, xs)
        assert (np.diff(qids) >= 0).all() and ys.min() >= 0
