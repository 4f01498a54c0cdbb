"""GPU parity tests proper: the CUDA path (through the C ABI) against the
oracle and the golden fixtures the real reference produced."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import make_model
from oracle import satrn, synth

import frx

pytestmark = pytest.mark.gpu

TAU = 1e-4        # margin below which a token may legitimately flip
LOGIT_TOL = 2e-4  # fp32 mode: max-abs logit error vs the reference (CPU fp32)


@pytest.fixture(scope="module")
def model0(ckpt0):
    m = make_model(ckpt0).cuda().eval()
    m.set_option("taps", 1)
    return m


def _nchw(tap):
    return tap.permute(0, 3, 1, 2).contiguous().cpu()


def test_trunk_and_encoder_taps_match_oracle(spec, ckpt0, model0):
    x = synth.synth_images(spec, 2, 0)
    taps = {}
    with torch.no_grad():
        mem_ref = satrn.encoder_forward(ckpt0, spec, x, taps=taps)
        mem = model0.encode(x.cuda())
    torch.cuda.synchronize()
    worst = {}
    for name, ref in taps.items():
        got = _nchw(model0.read_tap(name))
        assert got.shape == ref.shape, name
        scale = ref.abs().max().item() + 1e-6
        worst[name] = (got - ref).abs().max().item() / scale
    bad = {k: v for k, v in worst.items() if v > 1e-4}
    assert not bad, bad
    assert (mem.cpu() - mem_ref).abs().max().item() <= 1e-4 * (mem_ref.abs().max().item())


@pytest.mark.parametrize("seed", [0, 1])
def test_greedy_matches_reference_golden(spec, seed):
    g = load_golden(seed)
    sd = synth.synth_state_dict(spec, seed)
    model = make_model(sd).cuda().eval()
    b = g["tokens"].shape[0]
    x = synth.synth_images(spec, b, seed).cuda()
    with torch.no_grad():
        mem = model.encode(x)
        logits = model(x, satrn.expected_tokens(b).cuda(), False, 0.0)
        tokens = frx.decode(model, x, expected=satrn.expected_tokens(b).cuda(), method="greedy")
    assert logits.shape == (b, 231, 245) and logits.dtype == torch.float32 and logits.is_cuda
    assert tokens.shape == (b, 231) and tokens.dtype == torch.int64
    assert np.abs(mem.cpu().numpy() - g["memory"]).max() <= 1e-4 * np.abs(g["memory"]).max()
    ref_logits = torch.from_numpy(g["logits"])
    margins = satrn.min_margins(ref_logits)
    checked = 0
    for i in range(b):
        if margins[i] > TAU:
            assert np.array_equal(tokens[i].cpu().numpy(), g["tokens"][i]), i
            assert (logits[i].cpu() - ref_logits[i]).abs().max().item() <= LOGIT_TOL
            checked += 1
    assert checked >= b - 1
    # forced decoding: every sample and step, independent of token flips
    with torch.no_grad():
        fl, _ = model.greedy(x, 231, forced=torch.from_numpy(g["tokens"]))
    assert (fl.cpu() - ref_logits).abs().max().item() <= LOGIT_TOL


def test_decode_only_from_golden_memory(spec, ckpt0, model0):
    g = load_golden(0)
    mem = torch.from_numpy(g["memory"]).cuda()
    eng = model0.engine(mem.device, 4, 231)
    logits = torch.empty(4, 231, 245, device="cuda")
    tokens = torch.empty(4, 231, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    eng.h.call("frx_decode_greedy", mem.data_ptr(), 4, 231, logits.data_ptr(), tokens.data_ptr(), None, st)
    torch.cuda.synchronize()
    ref_logits = torch.from_numpy(g["logits"])
    ok = satrn.min_margins(ref_logits) > TAU
    assert torch.equal(tokens.cpu()[ok], torch.from_numpy(g["tokens"])[ok])
    assert (logits.cpu()[ok] - ref_logits[ok]).abs().max().item() <= LOGIT_TOL


def test_graph_and_eager_paths_agree(spec, ckpt0):
    x = synth.synth_images(spec, 3, 3).cuda()
    a = make_model(ckpt0).cuda().eval()
    b = make_model(ckpt0).cuda().eval()
    b.set_option("graphs", 0)
    with torch.no_grad():
        la, ta = a.greedy(x, 40)
        lb, tb = b.greedy(x, 40)
        la2, _ = a.greedy(x, 40)  # graph replay
    assert torch.equal(la, lb) and torch.equal(ta, tb) and torch.equal(la, la2)


def test_step_forward_api_matches_greedy(spec, ckpt0, model0):
    flags = frx.Flags(__import__("helpers").flags_dict()).get()
    dec = frx.EfficientSATRN_decoder(flags, __import__("helpers").Vocab(), None).cuda().eval()
    dec.load_state_dict({k: v for k, v in ckpt0.items() if k.startswith("decoder")}, strict=True)
    enc = frx.EfficientSATRN_encoder(flags, __import__("helpers").Vocab(), None).cuda().eval()
    enc.load_state_dict({k: v for k, v in ckpt0.items() if k.startswith("encoder")}, strict=True)
    x = synth.synth_images(spec, 2, 0).cuda()
    with torch.no_grad():
        src = enc(x)
        ref_logits, ref_tokens = model0.greedy(x, 12)
        dec.reset_status()
        target = torch.zeros(2, dtype=torch.int64, device="cuda")
        outs = []
        for t in range(12):
            o = dec.step_forward(src, target)
            assert o.shape == (2, 1, 245)
            target = o[:, -1].argmax(-1)
            outs.append(o)
    assert torch.equal(torch.cat(outs, 1), ref_logits)
    assert dec.step_idx == 12


def test_ragged_and_edge_batches(spec, ckpt0, model0):
    # batch 1 (the reference crashes there, SURVEY 3.1 -- we must not), odd batch, 1 step
    for b, steps in ((1, 5), (5, 1), (7, 9)):
        x = synth.synth_images(spec, b, 11).cuda()
        with torch.no_grad():
            lg, tk = model0.greedy(x, steps)
            ref = satrn.forward_greedy(ckpt0, spec, x.cpu(), steps)
        assert lg.shape == (b, steps, 245)
        assert (lg.cpu() - ref).abs().max().item() <= LOGIT_TOL
    with pytest.raises(RuntimeError):
        model0.engine(torch.device("cuda"), 2, 4).h.call("frx_decode_greedy", None, 0, 4, None, None, None, None)


def test_full_size_properties(spec, ckpt0):
    """BASELINE config size (B=256, 231 steps): properties that do not need the
    CPU oracle at full size -- run-to-run determinism, batch-slice invariance
    (each image's result is independent of its batch neighbours), and oracle
    agreement on a 4-image slice."""
    m = make_model(ckpt0, max_batch=256, max_steps=231).cuda().eval()
    x = synth.synth_images(spec, 256, 42).cuda()
    with torch.no_grad():
        l1, t1 = m.greedy(x, 231)
        l2, t2 = m.greedy(x, 231)
        ls, ts = m.greedy(x[100:104].contiguous(), 231)
    assert torch.equal(l1, l2) and torch.equal(t1, t2)
    assert torch.equal(l1[100:104], ls) and torch.equal(t1[100:104], ts)
    assert len({tuple(r.tolist()) for r in t1.cpu()}) > 32      # input-dependent outputs
    with torch.no_grad():
        ref = satrn.forward_greedy(ckpt0, spec, x[100:104].cpu(), 231)
    ok = satrn.min_margins(ref) > TAU
    assert torch.equal(ts.cpu()[ok], ref.argmax(-1)[ok])


def test_host_buffer_entry_point(spec, ckpt0, model0):
    x = synth.synth_images(spec, 4, 0).pin_memory()
    tok = model0.greedy_host(x, 231)
    with torch.no_grad():
        _, ref = model0.greedy(x.cuda(), 231)
    assert torch.equal(tok, ref.cpu())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_pipelined_host_entry_matches_synchronous_entry(ckpt0, spec, precision):
    """frx_forward_greedy_host_submit / _wait (copies on their own streams, the next batch's encoder on a third stream
    under the current batch's decode, two batches in flight) returns, batch by batch, exactly the tokens of the
    synchronous host entry."""
    model = make_model(ckpt0, precision=precision, max_batch=6, max_steps=16).cuda().eval()
    batches = [synth.synth_images(spec, 6, 50 + i).pin_memory() for i in range(5)]
    want = []
    eng = model.engine(torch.device("cuda", 0), 6, 16)
    st = torch.cuda.current_stream().cuda_stream
    for x in batches:
        t = torch.empty(6, 16, dtype=torch.int64).pin_memory()
        eng.h.call("frx_forward_greedy_host", x.data_ptr(), 6, 16, None, t.data_ptr(), st)
        want.append(t.clone())
    got = [torch.empty(6, 16, dtype=torch.int64).pin_memory() for _ in batches]
    for i, x in enumerate(batches):
        if i >= 2:
            model.wait_host(i % 2)
        model.submit_host(x, got[i], 16, i % 2)
    model.wait_host(0)
    model.wait_host(1)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    with pytest.raises(RuntimeError):          # a slot in flight cannot be re-submitted
        model.submit_host(batches[0], got[0], 16, 0)
        model.submit_host(batches[1], got[1], 16, 0)
    model.wait_host(0)
