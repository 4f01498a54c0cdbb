"""LiteSATRN (networks/LiteSATRN.py): oracle pinned against reference fixtures (CPU) and the CUDA path
against both (GPU)."""
import os

import numpy as np
import pytest
import torch

from helpers import ROOT, make_lite_model
from oracle import satrn, synth
from oracle.make_golden import LITE_SPEC, state_dict_digest

TAU, LOGIT_TOL = 1e-4, 2e-4


def _golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "litesatrn_seed0.npz"))


@pytest.fixture(scope="module")
def lite():
    spec = satrn.ModelSpec(**LITE_SPEC)
    return spec, synth.synth_state_dict(spec, 0, calib_batch=4)


def test_lite_layout_and_oracle_match_reference(lite):
    spec, sd = lite
    g = _golden()
    assert state_dict_digest(sd) == str(g["digest"])
    model = make_lite_model()
    assert list(model.state_dict().keys()) == list(sd.keys()) and len(sd) == 112
    assert sum(p.numel() for p in model.parameters()) == 2_633_077
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        mem = satrn.encoder_forward(sd, spec, synth.synth_images(spec, 3, 0))
        logits, tokens = satrn.decode_greedy(sd, spec, torch.from_numpy(g["memory"]), 120)
    assert np.abs(mem.numpy() - g["memory"]).max() <= 2e-5
    assert np.abs(logits.numpy() - g["logits"]).max() <= 1e-4
    assert np.array_equal(tokens.numpy(), g["tokens"])


@pytest.mark.gpu
def test_lite_cuda_path_matches_reference_golden(lite):
    spec, sd = lite
    g = _golden()
    model = make_lite_model(sd).cuda().eval()
    x = synth.synth_images(spec, 3, 0).cuda()
    with torch.no_grad():
        mem = model.encode(x)
        logits = model(x, satrn.expected_tokens(3, 119).cuda(), False, 0.0)
    assert mem.shape == (3, 128, 256)
    assert np.abs(mem.cpu().numpy() - g["memory"]).max() <= 1e-4 * np.abs(g["memory"]).max()
    ref = torch.from_numpy(g["logits"])
    ok = satrn.min_margins(ref) > TAU
    assert ok.sum() >= 2
    assert (logits.cpu()[ok] - ref[ok]).abs().max().item() <= LOGIT_TOL
    assert torch.equal(logits.cpu().argmax(-1)[ok], torch.from_numpy(g["tokens"])[ok])
    with pytest.raises(AttributeError):
        model.beam_search(x)


@pytest.mark.gpu
def test_lite_with_decoding_manager_matches_oracle(lite):
    """LiteSATRN.forward with a manager attached (LiteSATRN.py:516-543) vs the oracle's restatement of the rule
    machine (itself pinned against the reference's DecodingManager on EfficientSATRN, tests/test_oracle.py)."""
    from conftest import load_manager_golden
    from oracle import manager
    spec, sd = lite
    mg = load_manager_golden()
    rules = manager.Rules([str(t) for t in mg["vocab"]], mg["flags"], mg["limit"])
    model = make_lite_model(sd).cuda().eval()
    model.decoder.manager = rules.as_manager()
    x = synth.synth_images(spec, 3, 0)
    with torch.no_grad():
        mem = satrn.encoder_forward(sd, spec, x)
        ref_probs, ref_tokens = manager.decode_greedy_managed(sd, spec, mem, 60, rules)
        out = model(x.cuda(), satrn.expected_tokens(3, 59).cuda(), False, 0.0).cpu()
    assert out.shape == (3, 60, 245)
    agree = (out.argmax(-1) == ref_tokens).float().mean().item()
    assert agree >= 0.98, agree          # a near-tie may legitimately flip (fp32 on two devices)
    if agree == 1.0:
        assert (out - ref_probs).abs().max().item() <= 5e-5


LITE_BF16_MEM_TOL = 3e-2   # bf16 mode: ShallowCNN layers 1-3 on the tcgen05 GEMM with bf16 activations (memory, max rel)
LITE_BF16_REL_TOL = 8e-2   # + bf16 decoder (weights, KV cache): max |logit - ref| / max |ref| under forced decoding


@pytest.mark.gpu
def test_lite_bf16_mode_within_tolerance(lite):
    """bf16 mode: fused conv0 + pool, tcgen05 convs, and the greedy loop in the persistent cluster kernel compiled for
    LiteSATRN's decoder geometry (hidden 128, 4 heads, filter 512 -> clusters of 4 CTAs)."""
    spec, sd = lite
    g = _golden()
    model = make_lite_model(sd, precision="bf16").cuda().eval()
    x = synth.synth_images(spec, 3, 0).cuda()
    with torch.no_grad():
        mem = model.encode(x)
        steps = g["logits"].shape[1]
        logits, _ = model.greedy(x, steps, forced=torch.from_numpy(g["tokens"]))
        _, free = model.greedy(x, steps)
    rel_mem = np.abs(mem.cpu().numpy() - g["memory"]).max() / np.abs(g["memory"]).max()
    ref = torch.from_numpy(g["logits"])
    rel = ((logits.cpu() - ref).abs().max() / ref.abs().max()).item()
    agree_forced = (logits.cpu().argmax(-1) == ref.argmax(-1)).float().mean().item()
    agree_free = (free.cpu() == torch.from_numpy(g["tokens"])).float().mean().item()
    print("lite bf16: memory max rel %.4f, forced max rel logit error %.4f, per-step argmax agreement %.4f, "
          "free-running token agreement %.4f" % (rel_mem, rel, agree_forced, agree_free))
    assert rel_mem <= LITE_BF16_MEM_TOL
    assert rel <= LITE_BF16_REL_TOL
    assert agree_forced >= 0.9
