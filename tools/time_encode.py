"""Time the encoder alone (frx_encode: conv trunk + 2-D PE + encoder layers) with CUDA events: B = 256 synthetic images,
L2 flushed between passes.  A/B two builds by running it twice with FRX_LIBRARY=<path to the other libfrx.so>."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
model, _ = bench.build_model(a.precision, a.batch)
model = model.to(dev).eval()
x = bench.synthetic_images(a.batch, 0).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
with torch.no_grad():
    for _ in range(3):
        model.encode(x)
    torch.cuda.synchronize()
    ms = []
    for _ in range(a.reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model.encode(x)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
ms.sort()
print("%s encode B=%d: median %.3f ms, min %.3f ms (%s)" % (a.precision, a.batch, ms[len(ms) // 2], ms[0], os.environ.get("FRX_LIBRARY", "default library")))
