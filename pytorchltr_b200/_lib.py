"""ctypes binding of ``libltr_sm100.so`` (C ABI: ``include/ltr_sm100.h``).

There is no CPU fallback: if the shared library is missing, or a call fails, the
error is raised to the caller.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# LTR_SM100_LIB overrides the in-tree build (A/B testing of kernel variants)
LIB_PATH = os.environ.get("LTR_SM100_LIB") or os.path.join(_HERE, "csrc", "libltr_sm100.so")

# every symbol include/ltr_sm100.h declares
SYMBOLS = (
    "ltr_version", "ltr_strerror", "ltr_last_cuda_error", "ltr_pairwise_additive", "ltr_lambda",
    "ltr_listnet", "ltr_rank_metrics", "ltr_rank_by_score", "ltr_scale_rows",
    "ltr_host_workspace_bytes", "ltr_loss_host", "ltr_schedule_workspace_bytes",
    "ltr_pairwise_additive_ws", "ltr_lambda_ws", "ltr_host_workspace_dscores_offset",
    "ltr_linear_listnet_workspace_bytes", "ltr_linear_listnet", "ltr_linear_listnet_backward", "ltr_collate",
    "ltr_pbm_probabilities", "ltr_loss_host_ex", "ltr_scale_rows_host", "ltr_collate_sampled", "ltr_collate_sparse",
    "ltr_p2p_create", "ltr_p2p_connect", "ltr_p2p_allreduce_sum", "ltr_p2p_error", "ltr_p2p_destroy",
    "ltr_p2p_allreduce_vec", "ltr_mlp_backward_allreduce",
    "ltr_mlp_scores", "ltr_mlp_hz_pitch", "ltr_mlp_grad_len", "ltr_mlp_workspace_bytes", "ltr_mlp_backward",
)

ADD_HINGE, ADD_DCG_HINGE, ADD_LOGISTIC = 0, 1, 2
LAM_ARP1, LAM_ARP2, LAM_NDCG1, LAM_NDCG2 = 0, 1, 2, 3
METRIC_DCG, METRIC_NDCG, METRIC_ARP = 0, 1, 2
FAMILY_ADDITIVE, FAMILY_LAMBDA, FAMILY_LISTNET = 0, 1, 2
MAX_LIST_SIZE = 4096

_lock = threading.Lock()
_lib = None


class LtrError(RuntimeError):
    """A libltr_sm100 call returned a negative LTR_E* code."""


def _declare(lib):
    c_int, c_void_p, c_float, c_size_t = ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_size_t
    lib.ltr_version.restype = c_int
    lib.ltr_version.argtypes = []
    lib.ltr_strerror.restype = ctypes.c_char_p
    lib.ltr_strerror.argtypes = [c_int]
    lib.ltr_last_cuda_error.restype = c_int
    lib.ltr_last_cuda_error.argtypes = []
    lib.ltr_pairwise_additive.restype = c_int
    lib.ltr_pairwise_additive.argtypes = [c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                                          c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.ltr_lambda.restype = c_int
    lib.ltr_lambda.argtypes = [c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                               c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.ltr_schedule_workspace_bytes.restype = c_size_t
    lib.ltr_schedule_workspace_bytes.argtypes = [c_int]
    lib.ltr_pairwise_additive_ws.restype = c_int
    lib.ltr_pairwise_additive_ws.argtypes = [c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                                             c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_size_t, c_void_p]
    lib.ltr_lambda_ws.restype = c_int
    lib.ltr_lambda_ws.argtypes = [c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                  c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                  c_void_p]
    lib.ltr_listnet.restype = c_int
    lib.ltr_listnet.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p,
                                c_void_p, c_void_p, c_void_p]
    lib.ltr_rank_metrics.restype = c_int
    lib.ltr_rank_metrics.argtypes = [c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                     c_int, c_int, c_void_p, c_int, c_void_p]
    lib.ltr_rank_by_score.restype = c_int
    lib.ltr_rank_by_score.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.ltr_scale_rows.restype = c_int
    lib.ltr_scale_rows.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]
    lib.ltr_host_workspace_bytes.restype = c_size_t
    lib.ltr_host_workspace_bytes.argtypes = [c_int, c_int]
    lib.ltr_host_workspace_dscores_offset.restype = c_size_t
    lib.ltr_host_workspace_dscores_offset.argtypes = [c_int, c_int]
    lib.ltr_linear_listnet_workspace_bytes.restype = c_size_t
    lib.ltr_linear_listnet_workspace_bytes.argtypes = [c_int]
    lib.ltr_linear_listnet.restype = c_int
    lib.ltr_linear_listnet.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                                       c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.ltr_linear_listnet_backward.restype = c_int
    lib.ltr_linear_listnet_backward.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                                c_void_p, c_size_t, c_void_p]
    lib.ltr_pbm_probabilities.restype = c_int
    lib.ltr_pbm_probabilities.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int,
                                          c_float, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.ltr_collate.restype = c_int
    lib.ltr_collate.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p]
    lib.ltr_p2p_create.restype = c_int
    lib.ltr_p2p_create.argtypes = [c_int, c_int, ctypes.POINTER(c_void_p), c_void_p]
    lib.ltr_p2p_connect.restype = c_int
    lib.ltr_p2p_connect.argtypes = [c_void_p, c_void_p]
    lib.ltr_p2p_allreduce_sum.restype = c_int
    lib.ltr_p2p_allreduce_sum.argtypes = [c_void_p, c_void_p, c_int, c_void_p]
    lib.ltr_p2p_allreduce_vec.restype = c_int
    lib.ltr_p2p_allreduce_vec.argtypes = [c_void_p, c_void_p, ctypes.c_longlong, c_void_p]
    lib.ltr_p2p_error.restype = c_int
    lib.ltr_p2p_error.argtypes = [c_void_p]
    lib.ltr_p2p_destroy.restype = None
    lib.ltr_p2p_destroy.argtypes = [c_void_p]
    lib.ltr_mlp_scores.restype = c_int
    lib.ltr_mlp_scores.argtypes = [c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                   c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.ltr_mlp_hz_pitch.restype = c_int
    lib.ltr_mlp_hz_pitch.argtypes = [c_int, c_int]
    lib.ltr_mlp_grad_len.restype = c_size_t
    lib.ltr_mlp_grad_len.argtypes = [c_int, c_int, c_int]
    lib.ltr_mlp_workspace_bytes.restype = c_size_t
    lib.ltr_mlp_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.ltr_mlp_backward.restype = c_int
    lib.ltr_mlp_backward.argtypes = [c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                     c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                     c_void_p]
    lib.ltr_mlp_backward_allreduce.restype = c_int
    lib.ltr_mlp_backward_allreduce.argtypes = [c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                               c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_void_p, c_size_t, c_void_p, c_void_p]
    lib.ltr_collate_sampled.restype = c_int
    lib.ltr_collate_sampled.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.ltr_collate_sparse.restype = c_int
    lib.ltr_collate_sparse.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                       c_void_p, c_int, c_int, ctypes.c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p]
    lib.ltr_loss_host.restype = c_int
    lib.ltr_loss_host.argtypes = [c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float,
                                  c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.ltr_loss_host_ex.restype = c_int
    lib.ltr_loss_host_ex.argtypes = [c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float,
                                     c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]
    lib.ltr_scale_rows_host.restype = c_int
    lib.ltr_scale_rows_host.argtypes = [c_float, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]


def lib():
    """Loads the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
                        "`python -m pytorchltr_b200.build` (needs nvcc); pytorchltr_b200 has no "
                        "CPU fallback.")
                handle = ctypes.CDLL(LIB_PATH)
                missing = [s for s in SYMBOLS if not hasattr(handle, s)]
                if missing:
                    raise ImportError(f"{LIB_PATH} does not export {missing}")
                _declare(handle)
                _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        msg = lib().ltr_strerror(rc).decode()
        raise LtrError(f"libltr_sm100: {msg} (code {rc})")


# ---- device / stream plumbing for the callers of the library ---------------------------------------------------------
# torch.cuda.current_stream() builds a Stream object and torch.cuda.device() always switches twice: together ~20 us
# per call on the host, which is most of what a small-batch step costs.  These two do the same with one raw query
# each.
def raw_stream(device) -> int:
    """``cudaStream_t`` (as an int) of torch's current stream on ``device``."""
    import torch
    idx = device.index if device.index is not None else torch.cuda.current_device()
    try:
        return torch._C._cuda_getCurrentRawStream(idx)
    except AttributeError:                                # older / newer torch without the raw query
        return torch.cuda.current_stream(device).cuda_stream


class on_device:
    """``with on_device(dev):`` -- ``dev`` is the current CUDA device inside the block; switches (and switches back)
    only when it is not already."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index
        self.prev = -1

    def __enter__(self):
        import torch
        if self.idx is not None:
            cur = torch.cuda.current_device()
            if cur != self.idx:
                self.prev = cur
                torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev >= 0:
            import torch
            torch.cuda.set_device(self.prev)
        return False
