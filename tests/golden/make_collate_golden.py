"""Generates tests/golden/collate_ref.npz by running the UNMODIFIED reference collate_fn
(SVMRankDataset.collate_fn, datasets/svmrank/svmrank.py:135-205) with the default ListSampler on a
seeded ragged dataset.  Run in the build container only:

    python tests/golden/make_collate_golden.py
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("LTR_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

# svmrank.py imports the package's Cython SVMrank text parser at module level; it is not built in
# this container and collate_fn does not use it: stub the compiled module, leave the Python untouched
import types  # noqa: E402

_stub = types.ModuleType("pytorchltr.datasets.svmrank.parser.svmrank_parser")
_stub.parse_svmrank_file = None
sys.modules["pytorchltr.datasets.svmrank.parser.svmrank_parser"] = _stub

from pytorchltr.datasets.list_sampler import ListSampler  # noqa: E402
from pytorchltr.datasets.svmrank.svmrank import SVMRankDataset, SVMRankItem  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    g = torch.Generator().manual_seed(7)
    Q, F = 23, 12
    counts = torch.randint(1, 60, (Q,), generator=g)
    counts[3] = 1
    counts[5] = 59
    items = []
    for q in range(Q):
        n = int(counts[q])
        items.append(SVMRankItem(torch.randn(n, F, generator=g), torch.randint(0, 5, (n,), generator=g), n,
                                 100 + 3 * q, False))
    out = {"counts": counts.numpy(),
           "features": torch.cat([it.features for it in items]).numpy(),
           "relevance": torch.cat([it.relevance for it in items]).numpy(),
           "qids": np.array([it.qid for it in items], dtype=np.int64)}
    batches = {"all": list(range(Q)), "some": [5, 3, 17, 0, 22, 9], "one": [3], "rep": [4, 4, 11]}
    for mls in (None, 7, 40):
        fn = SVMRankDataset.collate_fn(ListSampler(mls))
        for name, idx in batches.items():
            b = fn([items[i] for i in idx])
            key = f"{name}_{mls}"
            out[f"{key}_idx"] = np.array(idx, dtype=np.int64)
            out[f"{key}_features"] = b.features.numpy()
            out[f"{key}_relevance"] = b.relevance.numpy()
            out[f"{key}_n"] = b.n.numpy()
            out[f"{key}_qid"] = b.qid.numpy()
    np.savez_compressed(os.path.join(HERE, "collate_ref.npz"), **out)
    print("wrote collate_ref.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
