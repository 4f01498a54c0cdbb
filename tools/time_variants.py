"""Greedy-inference timings of the LiteSATRN and SwinTRN variants (configs[4] of BASELINE.json): encode + 231-step greedy
decode through the public forward(), CUDA events, 3 timed passes after 2 warm-ups."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from helpers import make_lite_model, make_swin_model  # noqa: E402
from oracle import satrn, swin, synth  # noqa: E402
from oracle.make_golden import LITE_SPEC  # noqa: E402


def timed(model, x, steps):
    exp = satrn.expected_tokens(x.size(0), steps - 1).cuda()
    for _ in range(2):
        model(x, exp, False, 0.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        model(x, exp, False, 0.0)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3


lspec = satrn.ModelSpec(**LITE_SPEC)
lsd = synth.synth_state_dict(lspec, 0, calib_batch=4)
for prec in ("fp32", "bf16"):
    b = 256
    m = make_lite_model(lsd, precision=prec, max_batch=b, max_steps=231).cuda().eval()
    ms = timed(m, synth.synth_images(lspec, b, 0).cuda(), 231)
    print("LiteSATRN %s  B=%3d: %.1f ms per batch -> %.0f images/s" % (prec, b, ms, b / ms * 1e3), flush=True)
    del m
ck = swin.synth_state_dict(swin.swin_spec(), 0)
for prec in ("fp32", "bf16"):
    b = 16
    m = make_swin_model(ck, precision=prec, max_batch=b, max_steps=231).cuda().eval()
    ms = timed(m, swin.synth_images(b, 0).cuda(), 231)
    print("SwinTRN  %s  B=%3d: %.1f ms per batch -> %.0f images/s" % (prec, b, ms, b / ms * 1e3), flush=True)
    del m
