"""Ensemble decoding (utils/ensemble_utils.py:46-120, make_decoder_values) on the device.

The reference keeps every model's encoder output on the host (pickled), and per decode step calls
``model.step_forward(src_m, target)`` for every model, averages ``F.softmax`` of the logits, optionally passes the
average through ``DecodingManager.sift`` and takes the arg-max as the next target.  ``make_decoder_values`` here takes
the same decoder modules (``EfficientSATRN_decoder``) and the encoder memories as device tensors and runs the whole
loop inside the library (``frx_ensemble_decode``): the per-model logits, the averaged distribution and the next target
never leave HBM and there is no host synchronisation between steps."""
import ctypes

import torch

from . import decoding


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def make_decoder_values(models, memories, num_steps, manager=None):
    """``models``: decoder modules (EfficientSATRN_decoder or full EfficientSATRN), one per ensemble member;
    ``memories``: their encoder outputs [B, S, C] (CUDA tensors); ``num_steps`` = parser.max_sequence + 1 (:68);
    ``manager``: a reference-style DecodingManager (``.rules`` / ``.tokens``) or None.

    Returns (decoded_values [B, num_steps, V] -- the averaged (manager: masked) distributions the reference stacks at
    :112-115 -- and sequences [B, num_steps] int64, their arg-max (:117-118)).  Afterwards every model is back in its
    ``reset_status()`` state (:119-120)."""
    if not models or len(models) != len(memories):
        raise ValueError("one encoder memory per model")
    dev = memories[0].device
    if dev.type != "cuda":
        raise RuntimeError("frx runs on a CUDA device (B200) only; there is no CPU fallback")
    b = memories[0].size(0)
    engines, mems = [], []
    for m, mem in zip(models, memories):
        engines.append(m.engine(dev, b, max(num_steps, getattr(m, "_max_steps", 0) or 0)))
        mems.append(mem.detach().float().contiguous())
    h0 = engines[0].h
    if manager is not None and getattr(engines[0], "_rules_of", None) is not manager:
        flags, limit, ids6 = decoding.compile_decoding_rules(manager)
        n = len(flags)
        h0.call("frx_set_decoding_rules", (ctypes.c_int32 * n)(*flags), (ctypes.c_int32 * n)(*limit), n, (ctypes.c_int32 * 6)(*ids6))
        engines[0]._rules_of = manager
    v = models[0]._dims["num_classes"]
    probs = torch.empty(b, num_steps, v, device=dev)
    tokens = torch.empty(b, num_steps, dtype=torch.int64, device=dev)
    n = len(models)
    handles = (ctypes.c_void_p * n)(*[e.h.ptr for e in engines])
    mem_ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in mems])
    rc = h0.lib.frx_ensemble_decode(handles, n, mem_ptrs, b, num_steps, _ptr(probs), _ptr(tokens), 1 if manager is not None else 0,
                                    ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        raise RuntimeError("frx_ensemble_decode: " + h0.lib.frx_last_error(h0.ptr).decode())
    for m in models:
        if hasattr(m, "reset_status"):
            m.reset_status()
    return probs, tokens
