"""Time the SwinTRN encoder (Swin-B/384, 88.9 GFLOP per image) in both precision modes."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from helpers import make_swin_model  # noqa: E402
from oracle import swin  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ck = swin.synth_state_dict(swin.swin_spec(), 0)
for prec in ("fp32", "bf16"):
    m = make_swin_model(ck, precision=prec, max_batch=B, max_steps=8).cuda().eval()
    x = swin.synth_images(B, 0).cuda()
    for _ in range(2):
        m.encode(x)
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(5):
        m.encode(x)
    torch.cuda.synchronize()
    dt = (time.time() - t) / 5
    print("swin encoder %s: %.1f ms per %d images -> %.0f img/s, %.0f TFLOP/s"
          % (prec, dt * 1e3, B, B / dt, B * 88.9e9 / dt / 1e12))
    del m
