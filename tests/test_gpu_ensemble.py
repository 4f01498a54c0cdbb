"""Ensemble decoding (utils/ensemble_utils.py:46-120) on the device against the outputs of the REAL reference's
make_decoder_values on two seeded checkpoints (tests/golden/efficientsatrn_ensemble.npz, oracle/make_golden.py
--ensemble), without and with the DecodingManager."""
import os

import numpy as np
import pytest
import torch

import frx
from conftest import ROOT
from helpers import Vocab, flags_dict
from oracle import manager, satrn, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden", "efficientsatrn_ensemble.npz")


def _decoders(spec):
    flags = frx.Flags(flags_dict()).get()
    out = []
    for seed in (0, 1):
        sd = synth.synth_state_dict(spec, seed)
        dec = frx.EfficientSATRN_decoder(flags, Vocab(), {k: v for k, v in sd.items() if k.startswith("decoder.")},
                                         max_batch=4, max_steps=24).cuda().eval()
        out.append(dec)
    return out


@pytest.mark.parametrize("managed", [False, True])
def test_ensemble_matches_reference_golden(spec, managed):
    g = np.load(GOLD)
    models = _decoders(spec)
    mems = [torch.from_numpy(g["memory%d" % i]).cuda() for i in (0, 1)]
    mgr = manager.Rules([str(t) for t in g["vocab"]], g["flags"], g["limit"]).as_manager() if managed else None
    probs, tokens = frx.ensemble.make_decoder_values(models, mems, 24, mgr)
    torch.cuda.synchronize()
    tag = "managed" if managed else "plain"
    want_p, want_t = g["probs_" + tag], g["tokens_" + tag]
    err = np.abs(probs.cpu().numpy() - want_p).max()
    agree = (tokens.cpu().numpy() == want_t).mean()
    print("ensemble (%s): max |p - reference| = %.2e, token agreement %.4f" % (tag, err, agree))
    assert probs.shape == (3, 24, 245)
    assert err <= 2e-5
    assert agree == 1.0
    assert torch.equal(probs.argmax(-1), tokens)
    if managed:
        assert np.array_equal(probs.cpu().numpy() == 0, want_p == 0)     # the rule mask is exact
    for m in models:
        assert m.step_idx == 0                                             # reset_status (:119-120)


def test_ensemble_of_one_equals_the_single_model_softmax(spec, ckpt0):
    """A one-model ensemble is the softmax of that model's greedy logits (and its tokens the greedy tokens)."""
    g = np.load(GOLD)
    models = _decoders(spec)[:1]
    mem = torch.from_numpy(g["memory0"]).cuda()
    probs, tokens = frx.ensemble.make_decoder_values(models, [mem], 24, None)
    with torch.no_grad():
        lg, tok = satrn.decode_greedy(ckpt0, spec, mem.cpu(), 24)
    assert torch.equal(tokens.cpu(), tok)
    assert (probs.cpu() - torch.softmax(lg, -1)).abs().max().item() <= 2e-5
