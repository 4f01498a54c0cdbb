"""Time encode / decode phases (CUDA events inside the library) for several batch sizes."""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--batches", default="16,32,64,128,256")
ap.add_argument("--steps", type=int, default=231)
ap.add_argument("--dec-hpc", type=int, default=0, help="heads per CTA of the cluster decode kernel (1 or 2)")
ap.add_argument("--prof", type=int, default=1, help="0: no in-kernel stage profile (its clock reads serialise the profiled warp: use 0 for A/B timings)")
a = ap.parse_args()
NAMES = ["A qkv gemm", "self-attn", "cache rows", "attn wait", "N=32 gemms (B,C,D,G)", "other waits", "layernorm", "cross-attn", "E ffn0", "F ffn1", "argmax+embed"]
dev = torch.device("cuda", 0)
model, _ = bench.build_model(a.precision, 256)
model = model.to(dev).eval()
model.set_option("timing", 1)
model.set_option("prof", a.prof)
model.set_option("dec_hpc", a.dec_hpc)
ms3 = (ctypes.c_float * 4)()
for b in [int(x) for x in a.batches.split(",")]:
    x = bench.synthetic_images(b, 0).to(dev)
    eng = model.engine(dev, b, a.steps)
    tokens = torch.empty(b, a.steps, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        eng.h.call("frx_forward_greedy", x.data_ptr(), b, a.steps, None, tokens.data_ptr(), st)
    torch.cuda.synchronize()
    eng.h.lib.frx_last_timing(eng.h.ptr, ms3)
    print("B=%4d steps=%d  encode %.3f ms  decode %.3f ms (%.1f us/step)  total %.3f ms  -> %.0f img/s"
          % (b, a.steps, ms3[0], ms3[1], ms3[1] * 1e3 / a.steps, ms3[2], b / ms3[2] * 1e3), flush=True)
    if a.precision == "bf16" and a.prof:
        prof = (ctypes.c_int64 * 16)()
        eng.h.call("frx_read_prof", prof)
        tot = sum(prof[:len(NAMES)])
        print("   stage cycles (cluster 0, per step): " + ", ".join("%s %.0f" % (n, prof[i] / a.steps) for i, n in enumerate(NAMES)) + "  | total %.0f cyc/step" % (tot / a.steps), flush=True)
        if prof[12]:
            print("   self-attention of warp 0, cycles per step: K/V wait %.0f, ldmatrix + next request %.0f, scores + maximum %.0f, exp + P.V %.0f; %.2f blocks per call"
                  % (prof[11] / a.steps, prof[13] / a.steps, prof[14] / a.steps, prof[15] / a.steps, prof[12] / (3.0 * a.steps)), flush=True)
