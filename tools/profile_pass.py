"""One measured pass of the hot path bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off` (launch list or --set full captures)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=231)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--graphs", type=int, default=1)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    model, _ = bench.build_model(a.precision, a.batch)
    model = model.to(dev).eval()
    model.set_option("graphs", a.graphs)
    x = bench.synthetic_images(a.batch, 0).to(dev)
    with torch.no_grad():
        model.greedy(x, a.steps)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        model.greedy(x, a.steps)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print("profiled one pass: batch %d, %d decode steps, %s" % (a.batch, a.steps, a.precision))


if __name__ == "__main__":
    main()
