"""Compare the bf16 (tcgen05) encoder against the fp32 SIMT encoder and the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from helpers import make_model  # noqa: E402
from oracle import satrn, synth  # noqa: E402

spec = satrn.ModelSpec()
sd = synth.synth_state_dict(spec, 0)
x = synth.synth_images(spec, 8, 0).cuda()
m32 = make_model(sd).cuda().eval()
m16 = make_model(sd, precision="bf16").cuda().eval()
m16b = make_model(sd, precision="bf16").cuda().eval()
m16b.set_option("enc_fp32", 1)
m16.set_option("taps", 1)
with torch.no_grad():
    a = m32.encode(x)
    b = m16.encode(x)
    taps = {}
    ref = satrn.encoder_forward(sd, spec, x.cpu(), taps=taps).cuda()
    for name, t in taps.items():
        got = m16.read_tap(name).permute(0, 3, 1, 2).cpu()
        d = got - t
        print("tap %-16s rel L2 %.4f  max|d|/max %.4f" % (name, d.norm().item() / t.norm().item(), d.abs().max().item() / t.abs().max().item()))
    l32, t32 = m32.greedy(x, 231)
    l16, t16 = m16.greedy(x, 231, forced=t32)
    l16b, _ = m16b.greedy(x, 231, forced=t32)
    _, t16free = m16.greedy(x, 231)
def stats(name, got, want):
    d = (got - want).float()
    print("%-40s max|d|/max|ref| %.4f   rel L2 %.4f" % (name, d.abs().max().item() / want.abs().max().item(),
                                                         d.norm().item() / want.norm().item()))
stats("memory: fp32 path vs oracle", a, ref)
stats("memory: bf16 path vs oracle", b, ref)
stats("logits forced: bf16 enc+dec vs fp32", l16, l32)
stats("logits forced: fp32 enc + bf16 dec vs fp32", l16b, l32)
print("free-running token agreement bf16 vs fp32: %.4f" % (t16free == t32).float().mean().item())
print("per-step argmax agreement (forced): %.4f" % (l16.argmax(-1) == l32.argmax(-1)).float().mean().item())
