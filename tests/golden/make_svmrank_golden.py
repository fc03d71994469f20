#!/usr/bin/env python
"""Golden vectors for the SVMrank parser: synthetic files covering the grammar
(comments, CRLF, leading blanks, negative / fractional / exponent values, sparse and unordered columns,
no trailing newline, malformed inputs) and what the UNMODIFIED reference parser
(pytorchltr.datasets.svmrank.parser.svmrank_parser.parse_svmrank_file, built from /root/reference into
baseline/_ref) returns for each of them.  Run in the build container:

    python tests/golden/make_svmrank_golden.py        # writes tests/golden/svmrank/cases.npz
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
from pytorchltr.datasets.svmrank.parser.svmrank_parser import parse_svmrank_file  # noqa: E402


def synthetic_cases():
    rng = np.random.default_rng(20261017)
    cases = {}
    lines = []
    qid = 1
    for q in range(12):
        qid += int(rng.integers(1, 4))
        for _ in range(int(rng.integers(1, 9))):
            feats = " ".join("%d:%.6f" % (c, rng.random()) for c in range(1, 21))
            lines.append("%d qid:%d %s" % (rng.integers(0, 5), qid, feats))
    cases["dense"] = "\n".join(lines) + "\n"
    cases["no_trailing_newline"] = "\n".join(lines[:7])
    cases["comments_crlf"] = ("# header comment\r\n" + "\r\n".join(
        l + (" # docid = %d" % i if i % 2 else "") for i, l in enumerate(lines[:15])) + "\r\n#tail\r\n")
    sparse = []
    for i in range(40):
        cols = sorted(rng.choice(np.arange(3, 60), size=int(rng.integers(0, 7)), replace=False).tolist())
        if i == 5:
            cols = cols[::-1]                      # unordered columns
        toks = []
        for c in cols:
            kind = int(rng.integers(0, 5))
            v = float(rng.standard_normal()) * 10 ** int(rng.integers(-3, 4))
            if kind == 0:
                toks.append("%d:%d" % (c, int(abs(v)) % 1000))
            elif kind == 1:
                toks.append("%d:%.4f" % (c, v))
            elif kind == 2:
                toks.append("%d:%.3e" % (c, v))
            elif kind == 3:
                toks.append("%d:%.2E" % (c, v))
            else:
                toks.append("%d:-%.1f" % (c, abs(v)))
        lead = "  " if i % 7 == 0 else ""
        gap = "   " if i % 5 == 0 else " "
        sparse.append(lead + "%d%sqid:%d%s" % (rng.integers(0, 3), gap, 100 + i // 4,
                                               (gap + gap.join(toks)) if toks else " "))
    cases["sparse_mixed"] = "\n".join(sparse) + "\n"
    cases["zero_based_cols"] = "1 qid:3 0:0.5 2:1.5\n0 qid:3 1:2.5\n2 qid:4 0:1 1:2 2:3\n"
    cases["duplicate_cols"] = "1 qid:1 2:1.0 2:3.0 4:5.0\n0 qid:1 4:1.0\n"
    cases["big_ids"] = "3 qid:123456789012 7:0.25 9:-0.125\n0 qid:123456789013 8:1e-3\n".replace("1e-3", "1.0e-3")
    cases["empty"] = ""
    cases["only_comments"] = "# a\n#b\n"
    bad = {
        "bad_empty_line": "1 qid:1 1:0.5\n\n0 qid:1 1:0.25\n",
        "bad_letter": "1 qid:1 1:0.5x\n",
        "bad_no_qid": "1 1:0.5\n",
        "bad_exponent_without_fraction": "1 qid:1 1:5e3\n",
        "bad_negative_label": "-1 qid:1 1:0.5\n",
        "bad_tab": "1\tqid:1 1:0.5\n",
        "bad_missing_value": "1 qid:1 1:\n",
    }
    return cases, bad


def main():
    cases, bad = synthetic_cases()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, text in list(cases.items()) + list(bad.items()):
            path = os.path.join(tmp, name + ".txt")
            with open(path, "w", newline="") as f:
                f.write(text)
            out[name + "__text"] = np.frombuffer(text.encode("ascii"), dtype=np.uint8)
            try:
                xs, ys, qids = parse_svmrank_file(path)
                out[name + "__xs"], out[name + "__ys"], out[name + "__qids"] = np.array(xs), np.array(ys), np.array(qids)
                out[name + "__error"] = np.array(0)
            except ValueError:
                out[name + "__error"] = np.array(1)
            print(name, "error" if int(out[name + "__error"]) else out[name + "__xs"].shape)
    np.savez_compressed(os.path.join(HERE, "svmrank", "cases.npz"), **out)


if __name__ == "__main__":
    main()
