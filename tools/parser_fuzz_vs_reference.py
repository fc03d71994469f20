"""Differential fuzz of the SVMrank parser against the UNMODIFIED reference parser (the Cython build under baseline/_ref,
present in the build container only): random files mixing every token form of the grammar, one third of them with one
malformed token; values must be bit-identical and errors must coincide.  python tools/parser_fuzz_vs_reference.py [seed]
(Column ids stay small: the reference parser crashes on ids near 2^32.)"""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import test_svmrank_parser as T
sys.path.insert(0, "/root/repo/baseline/_ref")
from pytorchltr.datasets.svmrank.parser import parse_svmrank_file as ref_parse
from pytorchltr_b200.datasets.svmrank import parse_svmrank_file
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 11)
d = tempfile.mkdtemp()
ok = err = mism = 0
for trial in range(300):
    lines = []
    for _ in range(int(rng.integers(1, 30))):
        toks = [str(t) for t in rng.choice(T.HOSTILE_TOKENS, size=int(rng.integers(0, 10)))]
        toks += [f"{int(rng.integers(1, 300))}:{rng.random() * 10 ** int(rng.integers(-3, 6)):.{int(rng.integers(1, 9))}g}" for _ in range(int(rng.integers(1, 20)))]
        toks = [t for t in toks if "e" not in t.split(":")[1] or "." in t.split(":")[1]]
        rng.shuffle(toks)
        lines.append(T._line(rng, toks) + " ")      # a blank after the qid / last token: the reference needs it to store the qid
    if trial % 3 == 2:
        i = int(rng.integers(0, len(lines)))
        lines[i] = lines[i] + str(rng.choice(T.BROKEN_TOKENS))
    text = "\n".join(lines) + "\n"
    path = os.path.join(d, f"f{trial}.txt")
    open(path, "w").write(text)
    try:
        a = parse_svmrank_file(path, n_threads=2)
    except Exception as e:
        a = type(e).__name__
    try:
        b = ref_parse(path)
    except Exception as e:
        b = type(e).__name__
    if isinstance(a, str) or isinstance(b, str):
        err += 1
        if not (isinstance(a, str) and isinstance(b, str)):
            mism += 1; print("error mismatch", trial, a if isinstance(a, str) else "ok", b if isinstance(b, str) else "ok"); print(text[:300])
    else:
        ok += 1
        same = all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(a, b))
        if not same:
            mism += 1; print("value mismatch", trial)
print("ok", ok, "errors", err, "mismatches", mism)
