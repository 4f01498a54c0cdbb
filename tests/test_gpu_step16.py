"""Single decoder steps of a 16-bit handle on the persistent cluster kernel in "step mode" (one launch per step with a
per-image history length / ancestor chain / cache slot): the step_forward API, the ensemble loop and the best-first
search -- the paths that ran on the fp32 step kernels in round 1."""
import os

import numpy as np
import pytest
import torch

import frx
from conftest import ROOT, load_golden
from helpers import Vocab, flags_dict, make_model
from oracle import synth

pytestmark = pytest.mark.gpu


def _decode(model, mem, steps, forced=None):
    b = mem.size(0)
    eng = model.engine(mem.device, b, steps)
    logits = torch.empty(b, steps, 245, device="cuda")
    tokens = torch.empty(b, steps, dtype=torch.int64, device="cuda")
    f = forced.cuda().contiguous() if forced is not None else None
    eng.h.call("frx_decode_greedy", mem.data_ptr(), b, steps, logits.data_ptr(), tokens.data_ptr(),
               f.data_ptr() if f is not None else None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return logits, tokens


def test_step_mode_is_bitwise_the_greedy_kernel(ckpt0):
    """frx_decode_begin / frx_decode_step on a 16-bit handle: every step is one cluster-kernel launch that continues the
    bf16 K/V cache; fed the greedy tokens it must reproduce the all-steps-in-one-launch kernel bit for bit."""
    g = load_golden(0)
    mem = torch.from_numpy(g["memory"]).cuda()
    model = make_model(ckpt0, precision="bf16", max_batch=4, max_steps=40).cuda().eval()
    want_l, want_t = _decode(model, mem, 40)
    eng = model.engine(mem.device, 4, 40)
    st = torch.cuda.current_stream().cuda_stream
    launches0 = eng.launches
    eng.h.call("frx_decode_begin", mem.data_ptr(), 4, st)
    tgt = torch.zeros(4, dtype=torch.int64, device="cuda")          # <SOS>
    out = torch.empty(4, 245, device="cuda")
    for t in range(40):
        eng.h.call("frx_decode_step", tgt.data_ptr(), out.data_ptr(), st)
        assert torch.equal(out, want_l[:, t]), t
        tgt = want_t[:, t].contiguous()
    per_step = (eng.launches - launches0) / 40
    print("launches per step in step mode: %.1f" % per_step)
    assert per_step < 4


def test_beam_search_16bit_agrees_with_the_reference(ckpt0):
    """Best-first search with every node expansion on the cluster kernel (ancestor-chain K/V indirection): sequences
    against the real reference's fp32 outputs (a near-tie may flip a sample), and against the fp32 step kernels of the
    same handle (option step16 = 0)."""
    g = load_golden(0)
    mem = torch.from_numpy(g["memory"]).cuda()
    for bw in (4, 8):
        out = {}
        for step16 in (1, 0):
            model = make_model(ckpt0, precision="bf16").cuda().eval()
            model.set_option("step16", step16)
            eng = model.engine(mem.device, 4, 231)
            seq = torch.empty(4, 231, dtype=torch.int64, device="cuda")
            eng.h.call("frx_beam_search", mem.data_ptr(), 4, bw, 231, seq.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            out[step16] = seq.cpu().numpy()
        agree_ref = (out[1] == g["beam%d" % bw]).mean()
        agree_f32 = (out[1] == out[0]).mean()
        print("beam %d, 16-bit step mode: token agreement with the reference %.4f, with the fp32 step kernels %.4f" % (bw, agree_ref, agree_f32))
        assert (out[1][:, 0] == 0).all()
        assert agree_ref >= 0.90


def test_ensemble_with_16bit_decoders(spec):
    g = np.load(os.path.join(ROOT, "tests", "golden", "efficientsatrn_ensemble.npz"))
    flags = frx.Flags(flags_dict()).get()
    models = []
    for seed in (0, 1):
        sd = synth.synth_state_dict(spec, seed)
        models.append(frx.EfficientSATRN_decoder(flags, Vocab(), {k: v for k, v in sd.items() if k.startswith("decoder.")},
                                                 precision="bf16", max_batch=4, max_steps=24).cuda().eval())
    mems = [torch.from_numpy(g["memory%d" % i]).cuda() for i in (0, 1)]
    probs, tokens = frx.ensemble.make_decoder_values(models, mems, 24, None)
    torch.cuda.synchronize()
    err = np.abs(probs.cpu().numpy() - g["probs_plain"]).max()
    agree = (tokens.cpu().numpy() == g["tokens_plain"]).mean()
    print("ensemble of two 16-bit decoders: max |p - reference| = %.2e, token agreement %.4f" % (err, agree))
    # the average of two random-init models is a nearly flat distribution: its arg-max flips on differences far below the
    # probability error bound, so the token agreement is reported, not bounded tightly
    assert err <= 2e-2 and agree >= 0.5
    assert torch.equal(probs.argmax(-1), tokens)
