"""ssnt-tts-rust_b200 — B200 (sm_100a) backend for the ssnt-tts-rust hot path.

The product is ``libssnt_tts_c.so`` (CUDA kernels behind the reference's ``ssnt_tts_c``
C-ABI, see ``include/ssnt_tts_c.h``).  This module is the thin host-side mirror of the
reference's Python operator interface, ``ssnt-tts-tensorflow/ssnt_tts_tensorflow/__init__.py``:
the same function names, argument order and output order, over ``ctypes``.  Arguments may be

* numpy arrays (host buffers — what the reference's DEVICE_CPU ops pass; the library stages
  them to the GPU, runs the kernels and copies the results back), or
* torch CUDA tensors (device buffers — passed as raw pointers, nothing is copied, the call is
  asynchronous on torch's current stream).

There is no CPU implementation behind these functions: importing this module without the
built CUDA library raises, and every call needs a GPU.

The directory name is not a Python identifier; load it with ``importlib`` (see
``tests/conftest.py::load_product``) under the name ``ssnt_tts_rust_b200``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_bool, c_float, c_int, c_size_t, c_uint, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssnt_tts_c.so")

# Exported C symbols, in the order of include/ssnt_tts_c.h (tests check every one is present).
C_SYMBOLS = (
    "ssnt_tts_beam_search_decode",
    "ssnt_extract_best_beam_branch",
    "ssnt_tts_v2_beam_search_decode",
    "ssnt_order_beam_branch",
    "ssnt_upsample_source_indexes",
    "tone_latent_beam_search_decode",
    "tone_latent_levenshtein_edit_distance",
    "ssnt_tts_forward_backward_workspace_bytes",
    "ssnt_tts_forward_backward",
    "tone_latent_forward_backward_workspace_bytes",
    "tone_latent_forward_backward",
    "ssnt_tts_set_stream",
    "ssnt_tts_get_stream",
    "ssnt_tts_set_memory_space",
    "ssnt_tts_synchronize",
    "ssnt_tts_last_error",
    "ssnt_tts_set_fb_kernel",
    "ssnt_tts_get_fb_kernel_used",
    "ssnt_tts_set_tone_kernel",
    "ssnt_tts_get_tone_kernel_used",
    "ssnt_tts_fb_fallback_count",
    "ssnt_tts_debug_set_fb_stats",
    "ssnt_tts_backend",
    "ssnt_tts_debug_host_copy",
    "ssnt_tts_fill_i32",
    "ssnt_tts_v2_decode_loop",
    "ssnt_tts_forward_backward_logits_workspace_bytes",
    "ssnt_tts_forward_backward_logits",
    "tone_latent_decode_loop",
    "ssnt_tts_loss_exchange_export",
    "ssnt_tts_loss_exchange_connect",
    "ssnt_tts_loss_exchange_disconnect",
    "ssnt_tts_loss_allreduce",
)

ERR_V2_EMPTY_BEAM = 1
ERR_UPSAMPLE_LENGTH = 2
ERR_TONE_EMPTY_BEAM = 4
ERR_BAD_INDEX = 8

_lib = None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree with nvcc (sm_100a)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ssnt_b200_build", os.path.join(_HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force, verbose=verbose)


def lib() -> ctypes.CDLL:
    """The loaded C-ABI library.  Raises if it has not been built — there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python ssnt-tts-rust_b200/build.py` "
                "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.ssnt_tts_forward_backward_workspace_bytes.restype = c_size_t
        L.ssnt_tts_forward_backward_workspace_bytes.argtypes = [c_int, c_int, c_int]
        L.ssnt_tts_forward_backward_logits_workspace_bytes.restype = c_size_t
        L.ssnt_tts_forward_backward_logits_workspace_bytes.argtypes = [c_int, c_int, c_int]
        L.tone_latent_forward_backward_workspace_bytes.restype = c_size_t
        L.tone_latent_forward_backward_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
        L.ssnt_tts_get_stream.restype = c_void_p
        L.ssnt_tts_set_stream.argtypes = [c_void_p]
        L.ssnt_tts_last_error.restype = c_uint
        L.ssnt_tts_get_fb_kernel_used.restype = c_int
        L.ssnt_tts_get_tone_kernel_used.restype = c_int
        L.ssnt_tts_fb_fallback_count.restype = c_uint
        L.ssnt_tts_backend.restype = ctypes.c_char_p
        _lib = L
    return _lib


# ---- buffer plumbing ------------------------------------------------------------------------
def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _torch():
    import torch
    return torch


_NP = {"f32": np.float32, "i32": np.int32, "bool": np.bool_}


class _Call:
    """Per-call helper: normalises inputs to contiguous buffers of one memory space, allocates
    outputs in the same space, and hands raw pointers to ctypes."""

    def __init__(self, first):
        self.device = _is_torch(first) and first.is_cuda
        self.keep = []
        if self.device:
            torch = _torch()
            self.dev = first.device
            lib().ssnt_tts_set_stream(c_void_p(torch.cuda.current_stream(self.dev).cuda_stream))
        else:
            lib().ssnt_tts_set_stream(c_void_p(0))

    def inp(self, x, kind):
        if x is None:
            return None
        if self.device:
            torch = _torch()
            dt = {"f32": torch.float32, "i32": torch.int32, "bool": torch.bool}[kind]
            if not _is_torch(x):
                x = torch.as_tensor(np.ascontiguousarray(x, dtype=_NP[kind]), device=self.dev)
            x = x.to(device=self.dev, dtype=dt).contiguous()
            self.keep.append(x)
            return c_void_p(x.data_ptr())
        if _is_torch(x):
            x = x.detach().cpu().numpy()
        x = np.ascontiguousarray(x, dtype=_NP[kind])
        self.keep.append(x)
        return c_void_p(x.ctypes.data)

    def out(self, shape, kind, fill=None):
        if self.device:
            torch = _torch()
            dt = {"f32": torch.float32, "i32": torch.int32, "bool": torch.bool}[kind]
            o = (torch.empty(shape, dtype=dt, device=self.dev) if fill is None
                 else torch.full(shape, fill, dtype=dt, device=self.dev))
            return o, c_void_p(o.data_ptr())
        o = np.empty(shape, _NP[kind]) if fill is None else np.full(shape, fill, _NP[kind])
        return o, c_void_p(o.ctypes.data)


def set_fb_kernel(kind: int) -> None:
    """-1 auto, 0 generic block kernel, 1 log-domain warp/TMA cluster kernel, 2 block-float fused
    kernel, 4 block-float split-role kernel, 6 time-parallel block-float kernels (the default hot
    path), 8 warp-serial block-float kernels (the default for large batches); 3 / 5 / 7 / 9 = 2 / 4 / 6 / 8
    with the log-domain re-run forced (tests and benchmarks)."""
    lib().ssnt_tts_set_fb_kernel(c_int(kind))


def fb_kernel_used() -> int:
    return int(lib().ssnt_tts_get_fb_kernel_used())


def set_tone_kernel(kind: int) -> None:
    """Tone-latent lattice kernel for tests/benchmarks: -1 auto, 0 log domain, 1 split-role block-float, 2 warp-serial
    block-float, 3 = 2 with every utterance re-run in the log domain (include/ssnt_tts_c.h)."""
    lib().ssnt_tts_set_tone_kernel(c_int(kind))


def tone_kernel_used() -> int:
    return int(lib().ssnt_tts_get_tone_kernel_used())


def fb_fallback_count() -> int:
    """Cumulative number of utterances the block-float kernel re-ran in the log domain."""
    return int(lib().ssnt_tts_fb_fallback_count())


def synchronize() -> None:
    """Waits for the current stream; aborts the process if a device-side assert fired."""
    lib().ssnt_tts_synchronize()


def last_error() -> int:
    """Waits for the current stream, returns and clears the device-side assert bits."""
    return int(lib().ssnt_tts_last_error())


def backend() -> str:
    return lib().ssnt_tts_backend().decode()


# ---- the reference's seven operators (ssnt_tts_tensorflow/__init__.py) ---------------------------
def beam_search_decode(h, log_prob_history, is_finished, t, u, max_t, beam_width):
    """v1 Emit/Shift step, single batch (`__init__.py:8-21`; op SSNTBeamSearchDecode).
    Returns prediction, log_prob, next_t, next_u, is_finished, beam_branch — each [beam_width]."""
    c = _Call(h)
    W = int(beam_width)
    args = [c.inp(h, "f32"), c.inp(log_prob_history, "f32"), c.inp(is_finished, "bool"),
            c.inp(t, "i32"), c.inp(u, "i32")]
    pred, p_pred = c.out((W,), "i32", -1)
    lp, p_lp = c.out((W,), "f32")
    nt, p_nt = c.out((W,), "i32")
    nu, p_nu = c.out((W,), "i32")
    nf, p_nf = c.out((W,), "bool")
    bb, p_bb = c.out((W,), "i32")
    lib().ssnt_tts_beam_search_decode(*args, c_int(int(max_t)), c_int(W), p_pred, p_lp, p_nt, p_nu, p_nf, p_bb)
    return pred, lp, nt, nu, nf, bb


def extract_best_beam_branch(best_final_branch, beam_branch, t_history, beam_width):
    """`__init__.py:24-30`; op SSNTExtractBestBeamBranch.  beam_branch, t_history: [max_u, W]."""
    c = _Call(beam_branch)
    max_u = int(beam_branch.shape[0])
    a_bb, a_th = c.inp(beam_branch, "i32"), c.inp(t_history, "i32")
    ob, p_ob = c.out((max_u,), "i32")
    ot, p_ot = c.out((max_u,), "i32")
    lib().ssnt_extract_best_beam_branch(c_int(int(best_final_branch)), a_bb, a_th, c_int(int(beam_width)),
                                        c_int(max_u), p_ob, p_ot)
    return ob, ot


def ssnt_tts_v2_beam_search_decode(h, log_prob_history, is_finished, total_duration, duration_table, t, u,
                                   input_length, output_length, beam_width, duration_class_size,
                                   zero_duration_id, allow_skip, test_mode):
    """v2 duration-class step (`__init__.py:33-73`; op SSNTV2BeamSearchDecode).  As there,
    output_length is replaced by zeros in test mode (`:47`).  Returns prediction, log_prob, next_t,
    next_u, next_is_finished, next_total_duration, beam_branch — each [B, W]."""
    c = _Call(h)
    B, W = int(h.shape[0]), int(beam_width)
    if test_mode:  # in the caller's memory space: no host-to-device copy inside a device-pointer step
        output_length = (_torch().zeros(B, dtype=_torch().int32, device=c.dev) if c.device
                         else np.zeros(B, np.int32))
    args = [c.inp(h, "f32"), c.inp(log_prob_history, "f32"), c.inp(is_finished, "bool"),
            c.inp(total_duration, "i32"), c.inp(duration_table, "i32"), c.inp(t, "i32"), c.inp(u, "i32"),
            c.inp(input_length, "i32"), c.inp(output_length, "i32")]
    pred, p_pred = c.out((B, W), "i32", int(zero_duration_id))  # op pre-fill, _op.cc:149
    lp, p_lp = c.out((B, W), "f32")
    nt, p_nt = c.out((B, W), "i32")
    nu, p_nu = c.out((B, W), "i32")
    nf, p_nf = c.out((B, W), "bool")
    ntd, p_ntd = c.out((B, W), "i32")
    bb, p_bb = c.out((B, W), "i32")
    lib().ssnt_tts_v2_beam_search_decode(
        *args, c_int(B), c_int(W), c_int(int(duration_class_size)), c_int(int(zero_duration_id)),
        c_bool(bool(allow_skip)), c_bool(bool(test_mode)), p_pred, p_lp, p_nt, p_nu, p_nf, p_ntd, p_bb)
    return pred, lp, nt, nu, nf, ntd, bb


def order_beam_branch(final_branch, beam_branch, beam_width):
    """`__init__.py:76-82`; op SSNTOrderBeamBranch.  beam_branch [B,T,W] → [B,W,T]."""
    c = _Call(final_branch)
    B, T, W = (int(s) for s in beam_branch.shape)
    a_f, a_b = c.inp(final_branch, "i32"), c.inp(beam_branch, "i32")
    out, p_out = c.out((B, W, T), "i32")
    lib().ssnt_order_beam_branch(a_f, a_b, c_int(B), c_int(W), c_int(T), p_out)
    return out


def upsample_source_indexes(duration, output_length, out_of_range_source_index, beam_width, max_u=None):
    """`__init__.py:85-96`; op SSNTUpsampleSourceIndexes.  max_u defaults to max(output_length)
    (`:86`); the output is pre-filled with out_of_range_source_index (_op.cc:75)."""
    c = _Call(duration)
    B, W, T = (int(s) for s in duration.shape)
    if max_u is None:
        ol = output_length
        max_u = int(ol.max().item() if _is_torch(ol) else np.max(ol)) if B * W else 0
    a_d, a_l = c.inp(duration, "i32"), c.inp(output_length, "i32")
    out, p_out = c.out((B, W, int(max_u)), "i32", int(out_of_range_source_index))
    lib().ssnt_upsample_source_indexes(a_d, a_l, c_int(B), c_int(W), c_int(T), c_int(int(max_u)), p_out)
    return out


def tone_latent_beam_search_decode(h, log_prob_history, is_finished, t, u, input_length, beam_width,
                                   tone_class_size, empty_tone_id):
    """`__init__.py:99-127`; op ToneLatentBeamSearchDecode."""
    c = _Call(h)
    B, W = int(h.shape[0]), int(beam_width)
    args = [c.inp(h, "f32"), c.inp(log_prob_history, "f32"), c.inp(is_finished, "bool"), c.inp(t, "i32"),
            c.inp(u, "i32"), c.inp(input_length, "i32")]
    pred, p_pred = c.out((B, W), "i32", int(empty_tone_id))
    lp, p_lp = c.out((B, W), "f32")
    nt, p_nt = c.out((B, W), "i32")
    nu, p_nu = c.out((B, W), "i32")
    nf, p_nf = c.out((B, W), "bool")
    bb, p_bb = c.out((B, W), "i32")
    lib().tone_latent_beam_search_decode(*args, c_int(B), c_int(W), c_int(int(tone_class_size)),
                                         c_int(int(empty_tone_id)), p_pred, p_lp, p_nt, p_nu, p_nf, p_bb)
    return pred, lp, nt, nu, nf, bb


def levenshtein_edit_distance(a, b, a_lengths, b_lengths):
    """`__init__.py:130-134`; op ToneLatentLevenshteinEditDistance.  a, b: [B, max_length]."""
    c = _Call(a)
    B, L = int(a.shape[0]), int(a.shape[1])
    args = [c.inp(a, "i32"), c.inp(b, "i32"), c.inp(a_lengths, "i32"), c.inp(b_lengths, "i32")]
    out, p_out = c.out((B,), "i32")
    lib().tone_latent_levenshtein_edit_distance(*args, c_int(B), c_int(L), p_out)
    return out


# ---- whole-loop decoding (new) ---------------------------------------------------
def ssnt_tts_v2_decode_loop(h, duration_table, input_length, output_length, beam_width, duration_class_size,
                            zero_duration_id, allow_skip, test_mode, max_u, out_of_range_source_index,
                            log_prob_history=None, is_finished=None, total_duration=None, t=None, u=None):
    """Every v2 step of `ssnt_tts_v2_beam_search_decode` (`__init__.py:33-73`) for h[B, steps, W, D] in one launch,
    then `order_beam_branch` (final_branch = 0..W-1), the durations along each branch and
    `upsample_source_indexes` with output_length = the beams' final total durations (`__init__.py:76-96`).
    As in the per-step wrapper, output_length is replaced by zeros in test mode.  Returns a dict."""
    c = _Call(h)
    B, S, W = int(h.shape[0]), int(h.shape[1]), int(beam_width)
    if test_mode:
        output_length = (_torch().zeros(B, dtype=_torch().int32, device=c.dev) if c.device else np.zeros(B, np.int32))
    args = [c.inp(h, "f32"), c.inp(duration_table, "i32"), c.inp(input_length, "i32"), c.inp(output_length, "i32"),
            c.inp(log_prob_history, "f32"), c.inp(is_finished, "bool"), c.inp(total_duration, "i32"),
            c.inp(t, "i32"), c.inp(u, "i32")]
    ph, p_ph = c.out((B, S, W), "i32", int(zero_duration_id))
    bh, p_bh = c.out((B, S, W), "i32")
    lp, p_lp = c.out((B, W), "f32")
    ft, p_ft = c.out((B, W), "i32")
    fu, p_fu = c.out((B, W), "i32")
    ff, p_ff = c.out((B, W), "bool")
    ftd, p_ftd = c.out((B, W), "i32")
    ob, p_ob = c.out((B, W, S), "i32")
    du, p_du = c.out((B, W, S), "i32")
    # max_u = None: no upsampling (e.g. a loop resumed from a non-zero total_duration, whose source indexes would
    # not start at 0)
    up, p_up = (c.out((B, W, int(max_u)), "i32", int(out_of_range_source_index)) if max_u is not None else (None, None))
    max_u = 0 if max_u is None else max_u
    lib().ssnt_tts_v2_decode_loop(*args, c_int(B), c_int(S), c_int(W), c_int(int(duration_class_size)),
                                  c_int(int(zero_duration_id)), c_bool(bool(allow_skip)), c_bool(bool(test_mode)),
                                  c_int(int(max_u)), p_ph, p_bh, p_lp, p_ft, p_fu, p_ff, p_ftd, p_ob, p_du, p_up)
    return {"prediction_history": ph, "beam_branch_history": bh, "log_probs": lp, "t": ft, "u": fu, "is_finished": ff,
            "total_duration": ftd, "ordered_beam_branch": ob, "duration": du, "upsampled_source_indexes": up}


def tone_latent_decode_loop(h, input_length, beam_width, tone_class_size, empty_tone_id,
                            log_prob_history=None, is_finished=None, t=None, u=None):
    """Every step of `tone_latent_beam_search_decode` (`__init__.py:99-127`) for h[B, steps, W, K] in one launch, then
    the back-trace of every final beam and the tones along it.  Returns a dict."""
    c = _Call(h)
    B, S, W = int(h.shape[0]), int(h.shape[1]), int(beam_width)
    args = [c.inp(h, "f32"), c.inp(input_length, "i32"), c.inp(log_prob_history, "f32"), c.inp(is_finished, "bool"),
            c.inp(t, "i32"), c.inp(u, "i32")]
    ph, p_ph = c.out((B, S, W), "i32", int(empty_tone_id))
    bh, p_bh = c.out((B, S, W), "i32")
    lp, p_lp = c.out((B, W), "f32")
    ft, p_ft = c.out((B, W), "i32")
    fu, p_fu = c.out((B, W), "i32")
    ff, p_ff = c.out((B, W), "bool")
    ob, p_ob = c.out((B, W, S), "i32")
    ot, p_ot = c.out((B, W, S), "i32")
    lib().tone_latent_decode_loop(*args, c_int(B), c_int(S), c_int(W), c_int(int(tone_class_size)),
                                  c_int(int(empty_tone_id)), p_ph, p_bh, p_lp, p_ft, p_fu, p_ff, p_ob, p_ot)
    return {"prediction_history": ph, "beam_branch_history": bh, "log_probs": lp, "t": ft, "u": fu, "is_finished": ff,
            "ordered_beam_branch": ob, "ordered_tone": ot}


# ---- lattice forward-backward (new) -------------------------------------------------
def forward_backward_workspace_bytes(batch_size, max_t, max_u) -> int:
    return int(lib().ssnt_tts_forward_backward_workspace_bytes(batch_size, max_t, max_u))


def forward_backward(log_emit, log_shift, t_len=None, u_len=None, workspace=None, out=None):
    """Log-likelihood and gradients of the SSNT alignment lattice.

    log_emit, log_shift: [B, T, U] fp32.  Returns (log_likelihood[B], loss[1], grad_emit,
    grad_shift); gradients are d log_likelihood / d log-prob (posterior occupancies).
    ``workspace`` (torch CUDA uint8 tensor) and ``out`` (tuple of preallocated outputs) let a
    caller keep the call allocation-free."""
    c = _Call(log_emit)
    B, T, U = (int(s) for s in log_emit.shape)
    a_le, a_ls = c.inp(log_emit, "f32"), c.inp(log_shift, "f32")
    a_tl, a_ul = c.inp(t_len, "i32"), c.inp(u_len, "i32")
    if out is not None:
        ll, loss, ge, gs = out
        ptr = (lambda x: c_void_p(x.data_ptr())) if c.device else (lambda x: c_void_p(x.ctypes.data))
        p_ll, p_loss, p_ge, p_gs = ptr(ll), ptr(loss), ptr(ge), ptr(gs)
    else:
        ll, p_ll = c.out((B,), "f32")
        loss, p_loss = c.out((1,), "f32")
        ge, p_ge = c.out((B, T, U), "f32")
        gs, p_gs = c.out((B, T, U), "f32")
    ws_ptr, ws_bytes = c_void_p(0), 0
    if workspace is not None:
        ws_ptr, ws_bytes = c_void_p(workspace.data_ptr()), workspace.numel() * workspace.element_size()
    lib().ssnt_tts_forward_backward(a_le, a_ls, a_tl, a_ul, c_int(B), c_int(T), c_int(U), p_ll, p_loss,
                                    p_ge, p_gs, ws_ptr, c_size_t(ws_bytes))
    return ll, loss, ge, gs


def forward_backward_logits_workspace_bytes(batch_size, max_t, max_u) -> int:
    return int(lib().ssnt_tts_forward_backward_logits_workspace_bytes(batch_size, max_t, max_u))


def forward_backward_logits(logits, t_len=None, u_len=None, workspace=None, out=None):
    """The lattice on raw logits: log_emit = log sigmoid(z), log_shift = log sigmoid(-z) formed inside the kernels,
    gradient chained through them.  logits: [B, T, U] fp32.  Returns (log_likelihood[B], loss[1], grad_logits[B, T, U])
    with grad_logits = d log_likelihood / d z = grad_emit * sigmoid(-z) - grad_shift * sigmoid(z)."""
    c = _Call(logits)
    B, T, U = (int(s) for s in logits.shape)
    a_z = c.inp(logits, "f32")
    a_tl, a_ul = c.inp(t_len, "i32"), c.inp(u_len, "i32")
    if out is not None:
        ll, loss, gz = out
        ptr = (lambda x: c_void_p(x.data_ptr())) if c.device else (lambda x: c_void_p(x.ctypes.data))
        p_ll, p_loss, p_gz = ptr(ll), ptr(loss), ptr(gz)
    else:
        ll, p_ll = c.out((B,), "f32")
        loss, p_loss = c.out((1,), "f32")
        gz, p_gz = c.out((B, T, U), "f32")
    ws_ptr, ws_bytes = c_void_p(0), 0
    if workspace is not None:
        ws_ptr, ws_bytes = c_void_p(workspace.data_ptr()), workspace.numel() * workspace.element_size()
    lib().ssnt_tts_forward_backward_logits(a_z, a_tl, a_ul, c_int(B), c_int(T), c_int(U), p_ll, p_loss, p_gz,
                                           ws_ptr, c_size_t(ws_bytes))
    return ll, loss, gz


def tone_latent_forward_backward_workspace_bytes(batch_size, max_t, max_u, tone_class_size) -> int:
    return int(lib().tone_latent_forward_backward_workspace_bytes(batch_size, max_t, max_u, tone_class_size))


def tone_latent_forward_backward(log_emit, log_shift, log_tone, t_len=None, u_len=None, workspace=None, out=None):
    """Tone-latent marginalised lattice.  log_emit, log_shift: [B, T, U, K]; log_tone: [B, U, K].
    Returns (log_likelihood[B], loss[1], grad_emit, grad_shift, grad_tone).  ``workspace`` and ``out``
    as in forward_backward."""
    c = _Call(log_emit)
    B, T, U, K = (int(s) for s in log_emit.shape)
    a = [c.inp(log_emit, "f32"), c.inp(log_shift, "f32"), c.inp(log_tone, "f32"), c.inp(t_len, "i32"),
         c.inp(u_len, "i32")]
    if out is not None:
        ll, loss, ge, gs, gt = out
        ptr = (lambda x: c_void_p(x.data_ptr())) if c.device else (lambda x: c_void_p(x.ctypes.data))
        p_ll, p_loss, p_ge, p_gs, p_gt = ptr(ll), ptr(loss), ptr(ge), ptr(gs), ptr(gt)
    else:
        ll, p_ll = c.out((B,), "f32")
        loss, p_loss = c.out((1,), "f32")
        ge, p_ge = c.out((B, T, U, K), "f32")
        gs, p_gs = c.out((B, T, U, K), "f32")
        gt, p_gt = c.out((B, U, K), "f32")
    ws_ptr, ws_bytes = c_void_p(0), 0
    if workspace is not None:
        ws_ptr, ws_bytes = c_void_p(workspace.data_ptr()), workspace.numel() * workspace.element_size()
    lib().tone_latent_forward_backward(*a, c_int(B), c_int(T), c_int(U), c_int(K), p_ll, p_loss, p_ge, p_gs,
                                       p_gt, ws_ptr, c_size_t(ws_bytes))
    return ll, loss, ge, gs, gt


# ---- multi-GPU plumbing: batch sharding of independent utterances -------------------
def shard_range(batch_size: int, rank: int, world_size: int):
    """Contiguous shard [lo, hi) of the batch owned by `rank` (utterances are independent, as in
    the reference's `par_chunks` over the batch axis, src/v2.rs:227)."""
    per, rem = divmod(int(batch_size), int(world_size))
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def all_reduce_loss(loss, group=None):
    """Sum of the per-rank scalar losses: the only collective of the data-parallel path
    (torch.distributed, NCCL on GPUs / gloo in the CPU tests).  In place; returns `loss`."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
    return loss


def connect_loss_exchange(group=None) -> None:
    """Connects the NVLink loss exchange of the C library across the ranks of the initialised
    torch.distributed group (one process per GPU of one node): gathers every rank's CUDA IPC handle
    and opens the peers' slot buffers.  Afterwards each device-pointer ``forward_backward`` /
    ``tone_latent_forward_backward`` call also stores its loss into every rank's buffer from inside the
    kernel that reduces it; ``loss_allreduce()`` returns the sum of the latest call."""
    import torch.distributed as dist
    torch = _torch()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    buf = (ctypes.c_ubyte * 64)()
    lib().ssnt_tts_loss_exchange_export(c_int(world), buf)
    mine = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        mine = mine.cuda()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    blob = b"".join(bytes(t.cpu().tolist()) for t in gathered)
    handles = (ctypes.c_ubyte * (64 * world)).from_buffer_copy(blob)
    lib().ssnt_tts_loss_exchange_connect(c_int(rank), c_int(world), handles)
    dist.barrier(group)  # nobody stores into a peer before every rank has connected


def disconnect_loss_exchange() -> None:
    lib().ssnt_tts_loss_exchange_disconnect()


def loss_allreduce(out=None):
    """Sum over ranks of the latest call's loss (the same bits on every rank), as a 1-element CUDA tensor;
    asynchronous on torch's current stream."""
    torch = _torch()
    if out is None:
        out = torch.empty(1, dtype=torch.float32, device="cuda")
    lib().ssnt_tts_set_stream(c_void_p(torch.cuda.current_stream(out.device).cuda_stream))
    lib().ssnt_tts_loss_allreduce(c_void_p(out.data_ptr()))
    return out
