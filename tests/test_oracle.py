"""Pin the oracle (oracle/satrn.py) against the fixtures the REAL reference
produced (oracle/make_golden.py) -- CPU only, runs anywhere."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import satrn, synth
from oracle.make_golden import state_dict_digest

TAU = 1e-4          # margin below which a token may legitimately flip
LOGIT_TOL = 1e-4    # max-abs fp32 logit error, forced or free-running above margin


def test_state_dict_layout(spec, ckpt0):
    shapes = satrn.param_shapes(spec)
    assert len(shapes) == 923
    n_params = sum(int(np.prod(s)) for k, s in shapes.items()
                   if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert n_params == 27_221_141
    assert list(ckpt0) == list(shapes)
    trunk = sum(int(np.prod(s)) for k, s in shapes.items()
                if ".eff_block." in k and not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert trunk == 19_846_552                                  # timm tf_efficientnetv2_s .blocks


@pytest.mark.parametrize("seed", [0, 1])
def test_checkpoint_is_bit_reproducible(spec, seed):
    g = load_golden(seed)
    sd = synth.synth_state_dict(spec, seed, cache=False)
    assert state_dict_digest(sd) == str(g["digest"])


@pytest.mark.parametrize("seed", [0, 1])
def test_encoder_and_greedy_match_reference(spec, seed):
    g = load_golden(seed)
    sd = synth.synth_state_dict(spec, seed)
    b = g["memory"].shape[0]
    with torch.no_grad():
        mem = satrn.encoder_forward(sd, spec, synth.synth_images(spec, b, seed))
        assert np.abs(mem.numpy() - g["memory"]).max() <= 2e-5
        logits, tokens = satrn.decode_greedy(sd, spec, torch.from_numpy(g["memory"]), 231)
    ref_logits = torch.from_numpy(g["logits"])
    margins = satrn.min_margins(ref_logits)
    for i in range(b):
        if margins[i] > TAU:
            assert np.array_equal(tokens[i].numpy(), g["tokens"][i])
            assert (logits[i] - ref_logits[i]).abs().max() <= LOGIT_TOL
    assert (margins > TAU).sum() >= b - 1
    # forced decoding: every sample, every step
    with torch.no_grad():
        fl, _ = satrn.decode_greedy(sd, spec, torch.from_numpy(g["memory"]), 231,
                                    forced_tokens=torch.from_numpy(g["tokens"]))
    assert (fl - ref_logits).abs().max() <= LOGIT_TOL


def test_as_written_loop_matches_reference(spec, ckpt0):
    g = load_golden(0)
    with torch.no_grad():
        lg = satrn.decode_greedy_as_written(ckpt0, spec, torch.from_numpy(g["memory"][:2]), 40)
    assert np.abs(lg.numpy() - g["logits"][:2, :40]).max() <= LOGIT_TOL


@pytest.mark.parametrize("bw", [4, 8])
def test_beam_matches_reference(spec, ckpt0, bw):
    g = load_golden(0)
    with torch.no_grad():
        out = satrn.beam_search(ckpt0, spec, torch.from_numpy(g["memory"]), bw, 231)
    assert out.shape == (4, 231) and out.dtype == torch.int64
    assert np.array_equal(out.numpy(), g["beam%d" % bw])
    assert (out[:, 0] == satrn.SOS_ID).all()


def test_teacher_forced_matches_reference(spec, ckpt0):
    g = load_golden(0)
    with torch.no_grad():
        out = satrn.teacher_forced(ckpt0, spec, torch.from_numpy(g["memory"]),
                                   torch.from_numpy(g["tf_text"]))
    assert np.abs(out.numpy() - g["tf_logits"]).max() <= LOGIT_TOL


def test_expected_contract():
    e = satrn.expected_tokens(3)
    assert e.shape == (3, 232) and e[0, 0] == 0 and e[0, -1] == 1 and e[0, 1] == 158


def test_decoding_manager_matches_reference(spec, ckpt0):
    """Rule-constrained greedy decode (postprocessing.py:182-405, EfficientSATRN.py:536-564): tokens identical to
    the reference's, masked-softmax rows to 2e-6."""
    from conftest import load_manager_golden
    from oracle import manager
    g = load_manager_golden()
    rules = manager.Rules([str(t) for t in g["vocab"]], g["flags"], g["limit"])
    with torch.no_grad():
        mem = torch.from_numpy(load_golden(0)["memory"])
        probs, tokens = manager.decode_greedy_managed(ckpt0, spec, mem, 231, rules)
    assert np.array_equal(tokens.numpy(), g["tokens"])
    keep = g["probs_steps"]
    assert np.abs(probs[:, keep].numpy() - g["probs"]).max() <= 2e-6
    assert np.abs(probs.max(-1).values.numpy() - g["probs_max"]).max() <= 2e-6
