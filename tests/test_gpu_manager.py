"""Rule-constrained greedy decoding (the reference's DecodingManager, postprocessing/postprocessing.py:182-405,
attached by inference_single.py:80-95): CUDA path vs the reference's own outputs."""
import numpy as np
import pytest
import torch

from conftest import load_golden, load_manager_golden
from helpers import make_model
from oracle import manager, satrn, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rules():
    g = load_manager_golden()
    return manager.Rules([str(t) for t in g["vocab"]], g["flags"], g["limit"])


def _managed(model, mem, steps):
    b = mem.size(0)
    eng = model.engine(mem.device, b, steps)
    model._attach_manager(eng)
    probs = torch.empty(b, steps, 245, device="cuda")
    tokens = torch.empty(b, steps, dtype=torch.int64, device="cuda")
    eng.h.call("frx_decode_greedy_managed", mem.data_ptr(), b, steps, probs.data_ptr(), tokens.data_ptr(),
               torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return probs.cpu(), tokens.cpu()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_managed_decode_matches_reference_golden(ckpt0, rules, precision):
    g = load_manager_golden()
    model = make_model(ckpt0, precision=precision).cuda().eval()
    model.decoder.manager = rules.as_manager()
    mem = torch.from_numpy(load_golden(0)["memory"]).cuda()
    probs, tokens = _managed(model, mem, 231)
    # fp32: the step kernels + sift kernel; 16-bit: the persistent cluster kernel with the rules in its pick stage (a
    # free-running decode: a near-tie that flips one token changes the rest of that image's sequence)
    tol = 2e-5 if precision == "fp32" else 2e-2
    same = (tokens.numpy() == g["tokens"]).all(axis=1)
    agree = (tokens.numpy() == g["tokens"]).mean()
    print("managed decode (%s): token agreement %.4f, images identical %d / %d" % (precision, agree, same.sum(), len(same)))
    assert agree == 1.0 if precision == "fp32" else (agree >= 0.85 and same.sum() >= len(same) // 2)
    assert np.abs(probs[same][:, g["probs_steps"]].numpy() - g["probs"][same]).max() <= tol
    assert np.abs(probs[same].max(-1).values.numpy() - g["probs_max"][same]).max() <= tol
    assert torch.equal(probs.argmax(-1), tokens)          # the returned rows are the masked distributions of the run
    if agree < 1.0:
        return
    # the mask is exact: every class the reference zeroed is zero here and vice versa
    assert np.array_equal(probs[:, g["probs_steps"]].numpy() == 0, g["probs"] == 0)


def test_forward_with_manager_returns_masked_softmax(ckpt0, spec, rules):
    """model(input, expected, False, 0.0) with a manager attached (EfficientSATRN.py:536-564) vs the oracle, on an
    odd batch; also checks that the constrained tokens differ from the unconstrained ones (the rules bite)."""
    model = make_model(ckpt0).cuda().eval()
    x = synth.synth_images(spec, 3, 5)
    with torch.no_grad():
        mem = satrn.encoder_forward(ckpt0, spec, x)
        ref_probs, ref_tokens = manager.decode_greedy_managed(ckpt0, spec, mem, 40, rules)
        plain = model(x.cuda(), satrn.expected_tokens(3, 39).cuda(), False, 0.0).argmax(-1).cpu()
        model.decoder.manager = rules.as_manager()
        out = model(x.cuda(), satrn.expected_tokens(3, 39).cuda(), False, 0.0).cpu()
    assert out.shape == (3, 40, 245)
    assert torch.equal(out.argmax(-1), ref_tokens)
    assert (out - ref_probs).abs().max().item() <= 2e-5
    assert not torch.equal(plain, ref_tokens)
