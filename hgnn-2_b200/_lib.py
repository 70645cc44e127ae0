"""ctypes binding of libhgnn_b200.so (the C ABI of include/hgnn_b200.h).

There is NO fallback: if the shared library is missing the import fails, and every call checks the
status code and raises ``RuntimeError`` with ``hgnn_last_error()``.  Tensors must be CUDA,
contiguous and of the expected dtype - the host layer never silently computes on the CPU.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhgnn_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "hgnn_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; "
        "g.build()'` (nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

c_int, c_ll, c_float, c_void = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p

OP_IDENT, OP_DIAG, OP_CSR = 0, 1, 2
MAX_OPS = 8


class OpT(ctypes.Structure):
    _fields_ = [("kind", c_int), ("diag", c_void), ("rowptr", c_void), ("col", c_void),
                ("val", c_void), ("rng_rowptr", c_void), ("rng_id", c_void), ("rng_val", c_void),
                ("rng_lo", c_void), ("rng_hi", c_void), ("nnz", c_ll), ("rng_n", c_int)]


class SideT(ctypes.Structure):
    _fields_ = [("R", c_int), ("ops", ctypes.POINTER(OpT)), ("n_ops", c_int), ("Xs", c_void),
                ("Fs", c_int), ("p_rowptr", c_void), ("p_col", c_void), ("p_pm", c_void),
                ("p_pd", c_void), ("Xc", c_void), ("Fc", c_int), ("p_nnz", c_ll), ("roww", c_void), ("rowmap", c_void)]


class BnRefT(ctypes.Structure):
    _fields_ = [("acc", c_void), ("affine", c_void), ("weight", c_void), ("bias", c_void), ("n_rows", c_int)]


class SideBwdT(ctypes.Structure):
    _fields_ = [("gY", c_void), ("Z", c_void), ("Fg", c_int), ("relu_from", c_int), ("Rg", c_int),
                ("acc_f", c_void), ("acc_b", c_void), ("bn_weight", c_void),
                ("Wa", c_void), ("Ha", c_int), ("Wb", c_void), ("Hb", c_int), ("Cin", c_int),
                ("dW_bins", c_void), ("db_bins", c_void),
                ("R_self", c_int), ("ops_T", ctypes.POINTER(OpT)), ("n_ops", c_int), ("Xs", c_void),
                ("Fs", c_int), ("bn_self", BnRefT), ("gXs", c_void), ("accumulate_self", c_int),
                ("acc_b_self", c_void),
                ("R_cross", c_int), ("pt_rowptr", c_void), ("pt_col", c_void), ("pt_pm", c_void),
                ("pt_pd", c_void), ("Xc", c_void), ("Fc", c_int), ("bn_cross", BnRefT), ("gXc", c_void),
                ("accumulate_cross", c_int), ("acc_b_cross", c_void), ("skip_dw", c_int), ("pt_nnz", c_ll),
                ("rng_scratch", c_void), ("roww_self", c_void), ("roww_cross", c_void),
                ("rowmap_self", c_void), ("rowmap_cross", c_void)]


class ProgTensorT(ctypes.Structure):
    _fields_ = [("F", c_int), ("rows", c_int), ("bn_weight", c_int), ("bn_bias", c_int),
                ("acc_f", c_ll), ("acc_b", c_ll)]


class ProgSideT(ctypes.Structure):
    _fields_ = [("kind", c_int), ("src_self", c_int), ("src_cross", c_int), ("out", c_int),
                ("Wa", c_int), ("ba", c_int), ("Ha", c_int), ("Wb", c_int), ("bb", c_int), ("Hb", c_int),
                ("relu_from", c_int), ("dW_off", c_ll), ("db_off", c_ll)]


class ProgramT(ctypes.Structure):
    _fields_ = [("n_tensors", c_int), ("tensors", ctypes.POINTER(ProgTensorT)),
                ("n_sides", c_int), ("sides", ctypes.POINTER(ProgSideT)),
                ("dual", c_int), ("arena_doubles", c_ll),
                ("n_flat", c_int), ("red_off", c_void), ("red_nb", c_void), ("red_stride", c_void),
                ("red_cnt", c_void),
                ("n_bn", c_int), ("bn_acc_off", c_void), ("bn_F", c_void), ("bn_rows_kind", c_void),
                ("bn_run_off", c_void), ("momentum", c_float)]


class BatchT(ctypes.Structure):
    _fields_ = [("bs", c_int), ("Rn", c_int), ("Rm", c_int), ("n_ops", c_int),
                ("node_ops", ctypes.POINTER(OpT)), ("node_ops_T", ctypes.POINTER(OpT)),
                ("edge_ops", ctypes.POINTER(OpT)), ("edge_ops_T", ctypes.POINTER(OpT)),
                ("p_rowptr", c_void), ("p_col", c_void), ("p_pm", c_void), ("p_pd", c_void),
                ("pt_rowptr", c_void), ("pt_col", c_void), ("pt_pm", c_void), ("pt_pd", c_void),
                ("node_off", c_void), ("pad_n", c_void), ("p_nnz", c_ll),
                ("btc_rowptr", c_void), ("btc_col", c_void), ("btc_val", c_void),
                ("erow", c_void), ("ew", c_void), ("n_act", c_int), ("btc_nnz", c_ll), ("collapse_ok", c_int),
                ("mega_scratch", c_void)]


_P = c_void
_SIGS = {
    "hgnn_program_fwd": [ctypes.POINTER(ProgramT), ctypes.POINTER(BatchT), _P, _P, _P, _P, _P, _P, _P, _P],
    "hgnn_program_bwd": [ctypes.POINTER(ProgramT), ctypes.POINTER(BatchT), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_ll, _P],
    "hgnn_bn_running_update_k": [_P, _P, _P, _P, c_int, c_int, _P, c_int, c_float, _P, _P],
    "hgnn_host_pack_fill": [c_int, _P, c_int, c_int, _P, _P, c_int],
    "hgnn_host_fill_features": [c_int, _P, _P, c_int, c_ll, _P, c_ll, _P],
    "hgnn_pack_device_upload": [c_int, _P, c_int, c_int, _P, _P, _P, _P, _P],
    "hgnn_lg_side_fwd": [ctypes.POINTER(SideT), ctypes.POINTER(BnRefT), ctypes.POINTER(BnRefT), _P, _P, c_int,
                         _P, _P, c_int, c_int, _P, _P, _P, _P],
    "hgnn_lg_row4_eligible": [ctypes.POINTER(OpT), c_int, c_int, c_int, c_int],
    "hgnn_lg_wide_eligible": [c_int, c_int, c_int, c_int, c_int],
    "hgnn_lg_side_fits": [c_int, c_int, c_int, c_int],
    "hgnn_program_uses_collapse": [ctypes.POINTER(ProgramT), ctypes.POINTER(BatchT)],
    "hgnn_mega_set_trace": [_P],
    "hgnn_mega_grid_for": [c_int, c_int],
    "hgnn_lg_side_dw": [_P, _P, c_int, c_int, _P, _P, _P, _P, c_int, _P, _P, _P],
    "hgnn_lg_side_bwd": [ctypes.POINTER(SideBwdT), _P],
    "hgnn_debug_cta_times": [_P, c_int],
    "hgnn_debug_cta_phases": [_P, c_int],
    "hgnn_debug_ktrace": [_P, c_int, c_int],
    "hgnn_bins_reduce": [_P, _P, _P, _P, _P, c_int, _P, _P],
    "hgnn_bn_running_update": [_P, _P, _P, _P, _P, c_int, c_float, _P, _P],
    "hgnn_readout_bwd_prep": [_P, c_int, c_int, _P, _P, _P, _P, _P],
    "hgnn_pack_rows": [_P, c_int, c_int, c_int, _P, _P, _P],
    "hgnn_unpack_rows": [_P, c_int, c_int, c_int, _P, _P, _P, _P],
    "hgnn_dense_count_nnz": [_P, _P, c_ll, c_ll, c_ll, c_int, _P, _P, _P, _P],
    "hgnn_dense_fill_csr": [_P, _P, c_ll, c_ll, c_ll, c_int, _P, _P, _P, _P, _P, _P, _P],
    "hgnn_exclusive_scan_i32": [_P, _P, c_int, _P],
    "hgnn_csr_to_dense": [_P, _P, _P, c_int, _P, _P, _P, c_ll, c_ll, c_ll, _P],
    "hgnn_csr_row_sums": [_P, _P, c_int, _P, _P],
    "hgnn_fixup_offsets": [_P, _P, c_int, c_int, _P],
    "hgnn_spgemm_count_products": [c_int, _P, _P, _P, _P, _P],
    "hgnn_spgemm_expand": [c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "hgnn_spgemm_fill": [c_int, _P, _P, _P, _P, _P, _P, _P, c_int, _P],
    "hgnn_gmul_fwd": [ctypes.POINTER(OpT), c_int, c_int, c_int, _P, _P, _P],
    "hgnn_gmul_bwd": [ctypes.POINTER(OpT), c_int, c_int, c_int, _P, _P, _P],
    "hgnn_bn_stats": [_P, c_int, c_int, _P, _P, _P, _P, c_float, _P, _P, c_ll, _P],
    "hgnn_bn_stats_eval": [c_int, _P, _P, _P, _P, _P, _P],
    "hgnn_bn_apply": [_P, c_int, c_int, _P, _P, _P],
    "hgnn_bn_bwd_reduce": [_P, _P, c_int, c_int, _P, _P, c_int, _P, _P, _P, c_ll, _P],
    "hgnn_side_bwd_pre": [_P, _P, c_int, c_int, _P, c_int, _P, _P, _P, c_ll, _P],
    "hgnn_side_fwd": [ctypes.POINTER(SideT), _P, _P, c_int, _P, _P, c_int, c_int, _P, _P, _P, _P,
                      _P, c_float, _P, _P, c_ll, _P],
    "hgnn_side_bwd_gather": [ctypes.POINTER(OpT), c_int, c_int, _P, c_int, _P, c_int, _P, c_int, _P,
                             c_int, c_int, c_int, _P, c_int, _P, _P, _P, c_ll, _P],
    "hgnn_segment_sum": [_P, c_int, c_int, _P, _P, _P, _P, _P],
    "hgnn_segment_bcast": [_P, c_int, c_int, _P, _P, _P],
    "hgnn_ccn2_collapse6to3": [_P, c_int, c_int, _P, _P],
    "hgnn_ccn2_collapse6to3_bwd": [_P, c_int, c_int, _P, _P],
    "hgnn_ccn2_update_fwd": [c_int, c_int, _P, _P, _P, _P, c_int, _P, _P, c_int, _P, _P],
    "hgnn_ccn2_update_bwd": [c_int, c_int, _P, _P, _P, _P, c_int, _P, c_int, _P, _P, _P, _P, _P, _P, c_ll, _P],
    "hgnn_ccn1_update_fwd": [c_int, c_int, _P, _P, _P, c_int, _P, _P, c_int, _P, _P],
    "hgnn_ccn1_update_bwd": [c_int, c_int, _P, _P, _P, c_int, _P, c_int, _P, _P, _P, _P, _P, _P, c_ll, _P],
    "hgnn_adamax_step": [_P, _P, _P, _P, c_ll, c_float, c_float, c_float, c_float, c_float, _P, _P],
    "hgnn_p2p_alloc": [c_ll, _P, _P],
    "hgnn_p2p_open": [_P, _P],
    "hgnn_p2p_close": [_P],
    "hgnn_p2p_free": [_P],
    "hgnn_p2p_max_floats": [],
    "hgnn_p2p_allreduce_adamax": [_P, _P, _P, _P, c_int, c_float, c_float, c_float, c_float, c_float, _P, _P, c_int,
                                  c_int, c_ll, _P, _P],
}
EXPORTS = sorted(list(_SIGS) + ["hgnn_last_error", "hgnn_version", "hgnn_workspace_bytes", "hgnn_bins_for",
                                "hgnn_program_work_floats", "hgnn_program_launches", "hgnn_host_pack_n_keys",
                                "hgnn_host_pack_key", "hgnn_host_pack_layout", "hgnn_host_pack_last_ns",
                                "hgnn_pack_device_plan", "hgnn_lg_rng_scratch_bytes", "hgnn_program_rng_scratch_bytes",
                                "hgnn_p2p_buffer_bytes"])

for _name, _args in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = c_int
lib.hgnn_last_error.restype = ctypes.c_char_p
lib.hgnn_last_error.argtypes = []
lib.hgnn_version.restype = c_int
lib.hgnn_workspace_bytes.restype = c_ll
lib.hgnn_workspace_bytes.argtypes = [c_int]
lib.hgnn_bins_for.restype = c_int
lib.hgnn_bins_for.argtypes = [c_int]
lib.hgnn_program_work_floats.restype = c_ll
lib.hgnn_program_work_floats.argtypes = [ctypes.POINTER(ProgramT), c_int, c_int]
lib.hgnn_program_launches.restype = c_ll
lib.hgnn_program_launches.argtypes = []
lib.hgnn_host_pack_last_ns.restype = c_ll
lib.hgnn_host_pack_last_ns.argtypes = [c_int]
lib.hgnn_lg_rng_scratch_bytes.restype = c_ll
lib.hgnn_lg_rng_scratch_bytes.argtypes = [c_int]
lib.hgnn_program_rng_scratch_bytes.restype = c_ll
lib.hgnn_program_rng_scratch_bytes.argtypes = [ctypes.POINTER(ProgramT), ctypes.POINTER(BatchT)]
lib.hgnn_pack_device_plan.restype = c_ll
lib.hgnn_pack_device_plan.argtypes = [c_int, _P, c_int, c_int, _P, _P, _P]
lib.hgnn_host_pack_n_keys.restype = c_int
lib.hgnn_host_pack_n_keys.argtypes = []
lib.hgnn_host_pack_key.restype = ctypes.c_char_p
lib.hgnn_host_pack_key.argtypes = [c_int]
lib.hgnn_p2p_buffer_bytes.restype = c_ll
lib.hgnn_p2p_buffer_bytes.argtypes = [c_ll]
lib.hgnn_host_pack_layout.restype = c_ll
lib.hgnn_host_pack_layout.argtypes = [c_int, _P, c_int, c_int, _P]

# number of kernel launches issued through the C ABI (bench.py reports it as gpu_launches)
launch_count = 0
# optional per-call CUDA-event timing (bench.py): list of (name, tag, start_event, end_event)
timing = None
tag = ""


def call(name, *args):
    """Invoke one C-ABI entry point; raise on a non-zero status."""
    global launch_count
    if timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        timing.append((name, tag, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, lib.hgnn_last_error().decode()))
    launch_count += 1


def call_program(name, *args):
    """A hgnn_program_* entry point: the kernels it launches are counted on the C side
    (``hgnn_program_launches``), so the call itself is not added to ``launch_count``."""
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, lib.hgnn_last_error().decode()))


def total_launches():
    return launch_count + int(lib.hgnn_program_launches())


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("hgnn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("hgnn_b200: expected a CUDA tensor (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("hgnn_b200: expected a contiguous tensor")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError("hgnn_b200: expected dtype %s, got %s" % (dtype, t.dtype))
    return t.data_ptr()


def fptr(t):
    return ptr(t, torch.float32)


def iptr(t):
    return ptr(t, torch.int32)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def stream():
    """Raw handle of torch's current CUDA stream (the fast C accessors when this torch has them: the Python
    ``torch.cuda.current_stream()`` object costs ~14 us per call, and a step asks ~10 times)."""
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


_ws_cache = {}


def workspace(width, device):
    """Zero-initialised scratch for cross-CTA reductions (kernels reset their ticket counter, so
    one buffer per (device, stream) is reused by every call on that stream)."""
    need = int(lib.hgnn_workspace_bytes(int(width)))
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.zeros(max(need, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf.data_ptr(), buf.numel()


def make_ops(descs):
    """descs: list of ('ident',) | ('diag', vec) | ('csr', rowptr, col, val[, ranges]) -> (OpT array, n);
    ranges = (rng_rowptr, rng_id, rng_val, rng_lo, rng_hi) for run-length split operators."""
    n = len(descs)
    if n < 1 or n > MAX_OPS:
        raise RuntimeError("hgnn_b200: between 1 and %d operators are supported, got %d (J too large)"
                           % (MAX_OPS, n))
    arr = (OpT * n)()
    for i, d in enumerate(descs):
        if d[0] == "ident":
            arr[i].kind = OP_IDENT
        elif d[0] == "diag":
            arr[i].kind = OP_DIAG
            arr[i].diag = fptr(d[1])
        else:
            arr[i].kind = OP_CSR
            arr[i].rowptr, arr[i].col, arr[i].val = iptr(d[1]), iptr(d[2]), fptr(d[3])
            arr[i].nnz = d[2].numel()
            if len(d) > 4 and d[4] is not None and d[4][1].numel() > 0:
                r = d[4]
                arr[i].rng_rowptr, arr[i].rng_id, arr[i].rng_val = iptr(r[0]), iptr(r[1]), fptr(r[2])
                arr[i].rng_lo, arr[i].rng_hi = iptr(r[3]), iptr(r[4])
                arr[i].rng_n = r[3].numel()
    return arr, n
