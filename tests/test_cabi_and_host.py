"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares,
and the host-side mirror of the reference interface behaves (no compute without a GPU)."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions(header="ltr_sm100.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ltr_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_loads_and_exports_header_symbols():
    from pytorchltr_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    names = _header_functions()
    assert len(names) >= 11
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/ltr_sm100.h but not exported"
    assert sorted(_lib.SYMBOLS) == names
    lib = _lib.lib()
    assert lib.ltr_version() >= 105
    assert lib.ltr_strerror(0) == b"success"
    assert b"invalid" in lib.ltr_strerror(-1)
    assert lib.ltr_host_workspace_bytes(4, 8) >= 4 * 8 * 16


def test_parser_library_exports_header_symbols():
    from pytorchltr_b200 import build
    from pytorchltr_b200.datasets import svmrank
    path = build.build_parser()
    handle = ctypes.CDLL(path)
    names = _header_functions("ltr_svmrank.h")
    assert sorted(svmrank.SYMBOLS) == names
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/ltr_svmrank.h but not exported"


def test_argument_validation_needs_no_gpu():
    """Bad arguments are rejected before any CUDA call."""
    from pytorchltr_b200 import _lib
    lib = _lib.lib()
    assert lib.ltr_pairwise_additive(7, None, None, 8, None, 8, 1, 4, 1.0, None, None, None, None) == -1
    assert lib.ltr_lambda(0, None, None, 8, None, 8, 1, 0, 1.0, None, None, None, None, None) == -1
    assert lib.ltr_lambda(0, None, None, 8, None, 8, 1, _lib.MAX_LIST_SIZE + 1, 1.0, None, None, None,
                          None, None) == -2
    assert lib.ltr_rank_metrics(5, None, None, 8, None, 8, 1, 4, 1, 1, None, 1, None) == -1
    assert lib.ltr_scale_rows(None, 1, None, None, 3, 4, None) == -1
    assert lib.ltr_scale_rows(None, 2, None, None, 3, 4, None) == -1   # stride must be 0 or 1
    assert lib.ltr_scale_rows(None, 1, None, None, 0, 4, None) == 0   # empty batch is a no-op
    with pytest.raises(_lib.LtrError):
        _lib.check(-2)


def test_public_surface_mirrors_reference():
    import pytorchltr_b200 as p
    for name in ("PairwiseHingeLoss", "PairwiseDCGHingeLoss", "PairwiseLogisticLoss", "LambdaARPLoss1",
                 "LambdaARPLoss2", "LambdaNDCGLoss1", "LambdaNDCGLoss2", "ListNetLoss"):
        cls = getattr(p.loss, name)
        assert issubclass(cls, torch.nn.Module)
        params = inspect.signature(cls.forward).parameters
        assert list(params)[:4] == ["self", "scores", "relevance", "n"]
        # extensions beyond the reference's signature must be optional (loss_sum: distributed epilogue)
        assert all(q.default is not inspect.Parameter.empty for q in list(params.values())[4:])
    assert p.loss.PairwiseLogisticLoss(sigma=2.0).sigma == 2.0
    assert p.loss.LambdaNDCGLoss2().sigma == 1.0
    assert issubclass(p.loss.PairwiseDCGHingeLoss, p.loss.PairwiseHingeLoss)
    for fn, params in ((p.evaluation.ndcg, ["scores", "relevance", "n", "k", "exp"]),
                       (p.evaluation.dcg, ["scores", "relevance", "n", "k", "exp"]),
                       (p.evaluation.arp, ["scores", "relevance", "n"]),
                       (p.utils.rank_by_score, ["scores", "n", "generator"]),
                       (p.utils.mask_padded_values, ["xs", "n", "mask_value", "mutate"]),
                       (p.utils.tiebreak_argsort, ["x", "descending", "generator"]),
                       (p.utils.rank_by_plackettluce, ["scores", "n", "generator"]),
                       (p.utils.batch_pairs, ["x"])):
        assert list(inspect.signature(fn).parameters) == params


def test_shape_helpers_on_cpu():
    from pytorchltr_b200.utils import batch_pairs, mask_padded_values, tiebreak_argsort
    x = torch.tensor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    p = batch_pairs(x)
    assert p.shape == (2, 3, 3, 2)
    assert p[1, 0, 2, 0] == 4.0 and p[1, 0, 2, 1] == 6.0
    assert batch_pairs(x.reshape(2, 3, 1)).shape == (2, 3, 3, 2)
    m = mask_padded_values(x, torch.tensor([2, 0]))
    assert torch.isinf(m[0, 2]) and torch.isinf(m[1]).all() and m[0, 1] == 2.0 and x[0, 2] == 3.0
    mask_padded_values(x, torch.tensor([3, 1]), mask_value=0.0, mutate=True)
    assert x.tolist() == [[1.0, 2.0, 3.0], [4.0, 0.0, 0.0]]
    order = tiebreak_argsort(torch.tensor([[0.1, 0.9, 0.5]]), generator=torch.Generator().manual_seed(0))
    assert order.tolist() == [[1, 2, 0]]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_fails_loudly():
    import pytorchltr_b200 as p
    s = torch.zeros(2, 3, requires_grad=True)
    y = torch.zeros(2, 3, dtype=torch.long)
    n = torch.tensor([3, 2])
    with pytest.raises(RuntimeError, match="CUDA only"):
        p.loss.LambdaNDCGLoss2()(s, y, n)
    with pytest.raises(RuntimeError, match="CUDA only"):
        p.evaluation.ndcg(s, y, n, k=10)
    with pytest.raises(RuntimeError, match="CUDA only"):
        p.utils.rank_by_score(s, n)
    # the rows built beyond the loss path fail the same way: no silent CPU path anywhere
    with pytest.raises(RuntimeError, match="CUDA only"):
        p.fused.linear_listnet(torch.zeros(2, 3, 4), torch.zeros(4, requires_grad=True), None, y, n)
    with pytest.raises(RuntimeError, match="CUDA"):
        p.datasets.DeviceRankingDataset(torch.zeros(5, 4), torch.zeros(5, dtype=torch.long), torch.tensor([0, 2, 5]))
    with pytest.raises(RuntimeError, match="CUDA only"):
        p.click_simulation.simulate_perfect(torch.tensor([[0, 1, 2], [2, 1, 0]]), y, n)


def test_product_code_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "pytorchltr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "ltr_oracle" not in text, f
