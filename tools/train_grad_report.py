"""Per-tensor gradient error of the GPU training step against the train-step oracle (test infrastructure: imports oracle/).

    python tools/train_grad_report.py [--seed 0] [--batch 4] [--len 24] [--top 20 | --top 0 (every tensor, forward order)]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from helpers import make_model  # noqa: E402
from oracle import satrn, synth, train  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--len", type=int, default=24)
    ap.add_argument("--top", type=int, default=20)
    a = ap.parse_args()
    spec = satrn.ModelSpec()
    sd = synth.synth_state_dict(spec, a.seed)
    model = make_model(sd, max_batch=a.batch, max_steps=a.len).cuda().train()
    tr = train.Trainer(sd, spec)
    x, e = train.synth_batch(spec, a.batch, a.len, 10 * a.seed)
    loss, gn = model.train_step(x.cuda(), e.cuda())
    ref_loss, grads = tr.forward_backward(x, e)
    print("loss %.6f oracle %.6f  grad norm %.5f" % (loss.item(), ref_loss, gn.item()))
    rows = []
    for n, want in grads.items():
        got = model.read_grad(n).cpu()
        err = (got - want).norm().item() / max(want.norm().item(), 1e-12)
        if want.norm().item() < 1e-5:   # pure round-off tensors: absolute scale, as tests/test_gpu_train.py
            err = (got - want).norm().item() / 1e-3
        rows.append((err, want.norm().item(), (got - want).abs().max().item(), n))
    if a.top > 0:
        rows.sort(reverse=True)
        rows = rows[:a.top]
    for err, nrm, mx, n in rows:
        print("%-70s rel-L2 %.3e  |g| %.3e  max abs diff %.3e" % (n, err, nrm, mx))


if __name__ == "__main__":
    main()
