"""Mirror of ``pytorchltr.evaluation`` (reference: pytorchltr/evaluation/__init__.py:1-3).

``generate_pytrec_eval`` (host-side string formatting for pytrec_eval) is out of
scope of the accelerated path and is not provided.
"""
from pytorchltr_b200.evaluation.arp import arp  # noqa: F401
from pytorchltr_b200.evaluation.dcg import dcg  # noqa: F401
from pytorchltr_b200.evaluation.dcg import ndcg  # noqa: F401
