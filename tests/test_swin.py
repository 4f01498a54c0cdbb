"""SwinTRN (networks/SWIN.py): oracle pinned against reference fixtures (CPU) and the CUDA path against both (GPU)."""
import os

import numpy as np
import pytest
import torch

from helpers import ROOT, make_swin_model
from oracle import satrn, swin
from oracle.make_golden import state_dict_digest

LOGIT_TOL = 3e-4


def _golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "swin_seed0.npz"))


@pytest.fixture(scope="module")
def ckpt():
    return swin.synth_state_dict(swin.swin_spec(), 0)


def test_swin_layout_and_oracle_match_reference(ckpt):
    g = _golden()
    assert state_dict_digest(ckpt) == str(g["digest"])
    model = make_swin_model()
    sd = model.state_dict()
    assert list(sd.keys()) == list(ckpt.keys()) and len(sd) == 472
    for k in sd:
        assert sd[k].shape == ckpt[k].shape and sd[k].dtype == ckpt[k].dtype, k
    assert torch.equal(sd["encoder.layers.2.blocks.1.attn_mask"], ckpt["encoder.layers.2.blocks.1.attn_mask"])
    assert torch.equal(sd["encoder.layers.0.blocks.0.attn.relative_position_index"],
                       ckpt["encoder.layers.0.blocks.0.attn.relative_position_index"])
    model.load_state_dict(ckpt, strict=True)
    with torch.no_grad():
        mem = swin.encoder_forward(ckpt, swin.synth_images(2, 0))
        logits, tokens = satrn.decode_greedy(swin.decoder_view(ckpt), swin.swin_spec(), mem, 40)
    assert mem.shape == (2, 144, 1024)
    assert np.abs(mem[:, ::4].numpy() - g["memory_sub"]).max() <= 2e-5
    assert np.abs(logits.numpy() - g["logits"]).max() <= 1e-4
    assert np.array_equal(tokens.numpy(), g["tokens"])


@pytest.mark.gpu
def test_swin_cuda_path_matches_reference_golden(ckpt):
    g = _golden()
    model = make_swin_model(ckpt).cuda().eval()
    model.set_option("taps", 1)
    x = swin.synth_images(2, 0).cuda()
    taps = {}
    with torch.no_grad():
        mem = model.encode(x)
        logits = model(x, satrn.expected_tokens(2, 39).cuda(), False, 0.0)
        swin.encoder_forward(ckpt, x.cpu(), taps=taps)
    torch.cuda.synchronize()
    worst = {}
    for name, ref in taps.items():
        got = model.read_tap(name).flatten(1, 2).cpu()
        worst[name] = ((got - ref).abs().max() / (ref.abs().max() + 1e-6)).item()
    bad = {k: v for k, v in worst.items() if v > 2e-4}
    assert not bad, bad
    assert mem.shape == (2, 144, 1024)
    assert np.abs(mem.cpu()[:, ::4].numpy() - g["memory_sub"]).max() <= 2e-4 * np.abs(g["memory_sub"]).max()
    assert np.abs(logits.cpu().numpy() - g["logits"]).max() <= LOGIT_TOL
    assert np.array_equal(logits.cpu().argmax(-1).numpy(), g["tokens"])


SWIN_BF16_REL_TOL = 8e-2  # bf16 mode: 24 blocks x 4 linear layers on the tcgen05 GEMM (bf16 operands), fp32 residual
                          # stream; measured on the synthetic checkpoint: memory 0.064, logits 0.034, tokens identical


@pytest.mark.gpu
def test_swin_bf16_encoder_within_tolerance(ckpt):
    """bf16 mode: every linear layer of the Swin encoder on the tcgen05 GEMM; the decoder (hidden 512, 4 layers) keeps
    the fp32 step kernels.  Memory and logits vs the reference fixtures, tokens reported."""
    g = _golden()
    model = make_swin_model(ckpt, precision="bf16").cuda().eval()
    x = swin.synth_images(2, 0).cuda()
    with torch.no_grad():
        mem = model.encode(x)
        logits = model(x, satrn.expected_tokens(2, 39).cuda(), False, 0.0)
    torch.cuda.synchronize()
    ref = g["memory_sub"]
    rel_mem = np.abs(mem.cpu()[:, ::4].numpy() - ref).max() / np.abs(ref).max()
    rel_log = np.abs(logits.cpu().numpy() - g["logits"]).max() / np.abs(g["logits"]).max()
    agree = (logits.cpu().argmax(-1).numpy() == g["tokens"]).mean()
    print("swin bf16: memory max rel %.4f, logits max rel %.4f, token agreement %.3f" % (rel_mem, rel_log, agree))
    assert rel_mem <= SWIN_BF16_REL_TOL and rel_log <= SWIN_BF16_REL_TOL
    assert agree >= 0.9
