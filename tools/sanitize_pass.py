"""Small forwards of every bf16-mode path, meant to be run under compute-sanitizer (memcheck / racecheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from helpers import make_lite_model, make_model, make_swin_model  # noqa: E402
from oracle import satrn, swin, synth  # noqa: E402
from oracle.make_golden import LITE_SPEC  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
spec = satrn.ModelSpec()
if which in ("all", "eff"):
    sd = synth.synth_state_dict(spec, 0)
    m = make_model(sd, precision="bf16", max_batch=9, max_steps=6).cuda().eval()
    for hpc in (1, 2):
        m.set_option("dec_hpc", hpc)
        lg, tk = m.greedy(synth.synth_images(spec, 9, 0).cuda(), 6)
    torch.cuda.synchronize()
    print("EfficientSATRN bf16 ok", tk[0].tolist())
if which in ("all", "lite"):
    lspec = satrn.ModelSpec(**LITE_SPEC)
    lsd = synth.synth_state_dict(lspec, 0, calib_batch=4)
    m = make_lite_model(lsd, precision="bf16", max_batch=3, max_steps=6).cuda().eval()
    lg, tk = m.greedy(synth.synth_images(lspec, 3, 0).cuda(), 6)
    torch.cuda.synchronize()
    print("LiteSATRN bf16 ok", tk[0].tolist())
if which in ("all", "swin"):
    ck = swin.synth_state_dict(swin.swin_spec(), 0)
    m = make_swin_model(ck, precision="bf16", max_batch=1, max_steps=4).cuda().eval()
    mem = m.encode(swin.synth_images(1, 0).cuda())
    torch.cuda.synchronize()
    print("SwinTRN bf16 encoder ok", float(mem.abs().mean()))
