"""Data-parallel training step on N GPUs (run under torchrun, one process per GPU, NCCL): correctness of the bucketed
gradient all-reduce and its timing.  Test infrastructure (imports oracle/ for the synthetic checkpoint and batches).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/train_dp_check.py [--batch 4] [--len 24] [--time-batch 16]

Checks, with a different batch on every rank:
  1. the all-reduced flat gradient equals the sum of the ranks' local gradients (gathered separately);
  2. overlapped (bucket callbacks during the backward pass) and non-overlapped (CUDA graph + one all-reduce) steps leave
     the same parameters on a rank up to the atomics' summation order, and all ranks hold BIT-identical parameters after
     the update;
  3. prints ms per step for both modes at --time-batch x 231 positions, and the all-reduce alone.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from helpers import make_model  # noqa: E402
from oracle import satrn, synth, train  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--len", type=int, default=24)
    ap.add_argument("--time-batch", type=int, default=16)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    spec = satrn.ModelSpec()
    sd = synth.synth_state_dict(spec, 0)
    x, e = train.synth_batch(spec, a.batch, a.len, 100 + rank)
    x, e = x.to(dev), e.to(dev)
    names = ["decoder.generator.weight", "encoder.shallow_cnn.conv_stem.weight", "encoder.attention_layers.0.conv0.weight",
             "encoder.shallow_cnn.eff_block.4.3.conv_dw.weight"]

    def fresh():
        return make_model(sd, max_batch=max(a.batch, a.time_batch), max_steps=231).to(dev).train()

    # 1. local gradients (no reduction), gathered and summed by hand
    m0 = fresh()
    eng, tr = m0._trainer(dev, a.batch, a.len)
    sc = tr["scalars"]
    eng.h.call("frx_train_fwd_bwd", x.data_ptr(), e.data_ptr(), a.batch, a.len + 1, sc.data_ptr(), torch.cuda.current_stream().cuda_stream)
    local_g = tr["grads"].clone()
    want = local_g.clone()
    dist.all_reduce(want)
    results = {}
    for overlap in (True, False):
        m = fresh()
        loss, gn = m.train_step(x, e, overlap=overlap)
        got = m._engine._train["grads"].clone()
        err = (got - want).abs().max().item() / want.abs().max().item()
        m.sync_trained_weights()
        results[overlap] = {n: m.state_dict()[n].clone() for n in names}
        print("[rank %d] overlap=%s loss %.6f grad-norm(mean grad) %.5f  |allreduced - sum of locals| / max = %.2e"
              % (rank, overlap, loss.item(), gn.item(), err), flush=True)
        assert err < 1e-5, err
    for n in names:
        # two runs accumulate weight gradients with atomics in a different order; Adam's first step moves every element by
        # ~lr whatever the size of its gradient, so round-off-sized gradients may step in opposite directions
        assert (results[True][n] - results[False][n]).abs().max().item() <= 2.1 * 5e-4, n
        assert (results[True][n] - results[False][n]).abs().mean().item() <= 2e-5, n
        ref = results[True][n].clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, results[True][n]), "rank %d holds different %s after the update" % (rank, n)
    if rank == 0:
        print("parameters bit-identical across ranks; overlap modes agree to round-off", flush=True)
    # 3. timings
    B, L = a.time_batch, 231
    xb, eb = train.synth_batch(spec, B, L, 7 + rank)
    xb, eb = xb.to(dev), eb.to(dev)
    for overlap in (True, False):
        m = fresh()
        for _ in range(3):
            m.train_step(xb, eb, overlap=overlap)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            m.train_step(xb, eb, overlap=overlap)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("B=%d x %d positions per GPU, %d GPUs, overlap=%s: %.2f ms per step -> %.0f images/s"
                  % (B, L, world, overlap, ms.item(), B * world / ms.item() * 1e3), flush=True)
    g = m._engine._train["grads"]
    dist.all_reduce(g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dist.all_reduce(g)
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print("all-reduce of the %.1f MB gradient buffer alone: %.3f ms" % (g.numel() * 4 / 1e6, e0.elapsed_time(e1) / 5), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
