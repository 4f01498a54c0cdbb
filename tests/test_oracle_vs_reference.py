"""Live check of the oracle against the reference's own modules (only where
/root/reference exists, i.e. the build container)."""
import pytest
import torch

from oracle import ref_shim, satrn, synth

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


def test_live_reference_forward(spec, ckpt1):
    ref = ref_shim.load_reference()
    model = ref.networks.EfficientSATRN(ref_shim.reference_flags(), ref_shim.reference_vocab()).eval()
    model.load_state_dict(ckpt1, strict=True)
    x = synth.synth_images(spec, 2, 5)
    with torch.no_grad():
        out_ref = model(x, satrn.expected_tokens(2, 30), False, 0.0)
        mem_ref = model.encoder(x)
        mem = satrn.encoder_forward(ckpt1, spec, x)
        out = satrn.decode_greedy(ckpt1, spec, mem, 31)[0]
    assert (mem - mem_ref).abs().max() <= 1e-5
    assert (out - out_ref).abs().max() <= 1e-4
