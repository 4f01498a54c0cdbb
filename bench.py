#!/usr/bin/env python
"""bench.py -- EfficientSATRN greedy-decode images/sec (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A "step" is one pass of the hot path (encode + 231-step greedy decode) over one
synthetic batch of 256 images (128x256 grayscale, randn) per GPU.  `value` is
timed with CUDA events with the images already resident in HBM; `e2e` goes
through the host-buffer C-ABI entry (frx_forward_greedy_host): H2D of the pinned
images + compute + D2H of the tokens inside the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MAX_SEQUENCE = 230          # inference.py:29 -> expected has 232 ids -> 231 steps
STEPS_PER_IMAGE = MAX_SEQUENCE + 1
METRIC = "EfficientSATRN greedy-decode images/sec"
# SURVEY 8d / BASELINE.md: algorithmic work per image
DECODE_BYTES_PER_IMAGE = {"fp32": 210.5e6, "bf16": 105.4e6}
ENCODE_FLOP_PER_IMAGE = 3.700e9 + 0.207e9
TOTAL_FLOP_PER_IMAGE = 5.54e9


# the decode kernel the library launches for batches of more than 15 clusters (two heads per CTA, clusters of 4)
DECODE_KERNEL = "dec_cluster_bf16_kernel_p2"


def ncu_traffic(kernel, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    (profiles/ncu_traffic.json), or None when no capture matches this kernel / batch."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    for e in json.load(open(path)):
        if e["kernel"] == kernel and e["batch"] == batch:
            return e["dram_bytes_per_launch"]
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.

    The sampler process is started before the warm-up (nvidia-smi needs a few hundred ms to produce its first
    line on an 8-GPU box); every line is time-stamped on arrival and only the lines that fall inside
    [window_begin(), window_end()] are summarised."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def window_begin(self):
        self.t0 = time.monotonic()

    def window_end(self):
        self.t1 = time.monotonic()

    def stop(self):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if self.t0 is not None and not (self.t0 <= t <= self.t1 + 0.02):
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_images(batch, seed):
    import torch
    g = torch.Generator().manual_seed(2000 + seed)
    return torch.randn(batch, 1, 128, 256, generator=g)


def workload_config(batch, precision):
    return {"workload": "EfficientSATRN greedy inference, batch %d per GPU, 128x256x1 randn images, "
                        "max_sequence 230 (231 decode steps), random-init weights, image batch sharded "
                        "over ranks (no collective)" % batch,
            "batch_per_gpu": batch, "decode_steps": STEPS_PER_IMAGE, "precision": precision,
            "l2": "256 MiB buffer written between timed iterations (L2 flush); per-step working set "
                  "(KV cache + activations) also exceeds the 126 MB L2"}


def build_model(precision, batch):
    import frx
    from helpers import Vocab, flags_dict
    flags = frx.Flags(flags_dict()).get()
    dims = frx.layout.dims_from_flags(flags, 245)
    sd = frx.synthetic.synthetic_state_dict(dims, seed=0)
    model = frx.EfficientSATRN(flags, Vocab(), sd, None, precision=precision, max_batch=batch,
                               max_steps=STEPS_PER_IMAGE)
    return model, sd


def cpu_reference_throughput(sd, sample_batch, runs, as_written=True):
    """The reference's CPU algorithm (oracle port, op for op incl. the per-step
    K/V re-projection) on the host cores.  Returns (images/s, cores, tokens)."""
    import torch
    from oracle import satrn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = satrn.ModelSpec()
    x = synthetic_images(sample_batch, 0)
    with torch.no_grad():
        satrn.forward_greedy(sd, spec, x[:2], 4, as_written=as_written)  # warm-up
        t0 = time.perf_counter()
        for _ in range(runs):
            logits = satrn.forward_greedy(sd, spec, x, STEPS_PER_IMAGE, as_written=as_written)
        dt = time.perf_counter() - t0
    return sample_batch * runs / dt, cores, logits.argmax(-1), dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    import frx
    from helpers import flags_dict
    dims = frx.layout.dims_from_flags(frx.Flags(flags_dict()).get(), 245)
    sd = frx.synthetic.synthetic_state_dict(dims, seed=0)
    sample = args.cpu_sample
    for _ in range(args.warmup):
        pass  # the warm-up happens inside cpu_reference_throughput on a tiny batch
    ips, cores, _, dt = cpu_reference_throughput(sd, sample, max(1, args.steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / max(1, args.steps) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args.batch, args.precision),
                       reference_arm="the reference's CPU algorithm (oracle port, op for op, fp32) on the host cores; "
                                     "each step is a bounded sample of %d images of the same workload" % sample),
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d step(s) x %d images x 231 decode steps, torch %s CPU fp32, %d threads"
                                   % (max(1, args.steps), sample, torch.__version__, cores)},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_frx(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU for --impl frx (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B, T = args.batch, STEPS_PER_IMAGE
    model, sd = build_model(args.precision, B)
    model = model.to(dev).eval()
    model.set_option("timing", 1)
    images_host = synthetic_images(B, rank).pin_memory()
    images = images_host.to(dev)
    logits = torch.empty(B, T, 245, device=dev)
    tokens = torch.empty(B, T, dtype=torch.int64, device=dev)
    tokens_host = torch.empty(B, T, dtype=torch.int64).pin_memory()
    eng = model.engine(dev, B, T)
    stream = torch.cuda.current_stream(dev).cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_device():
        eng.h.call("frx_forward_greedy", images.data_ptr(), B, T, logits.data_ptr(), tokens.data_ptr(), stream)

    def step_host():
        eng.h.call("frx_forward_greedy_host", images_host.data_ptr(), B, T, None, tokens_host.data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    clocks = ClockSampler(local_rank).start()
    for _ in range(max(args.warmup, 1)):
        step_device()
    torch.cuda.synchronize(dev)
    import ctypes
    ms3 = (ctypes.c_float * 4)()

    # ---- timed region 1: device-resident inputs, CUDA events -------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = eng.launches
    enc_ms, dec_ms, kern_ms = [], [], []
    barrier()
    clocks.window_begin()
    for s, e in ev:
        flush.zero_()            # L2 flush between timed iterations (outside the event pair)
        s.record()
        step_device()
        e.record()
        e.synchronize()
        eng.h.lib.frx_last_timing(eng.h.ptr, ms3)
        enc_ms.append(ms3[0]); dec_ms.append(ms3[1]); kern_ms.append(ms3[3])
    barrier()
    clocks.window_end()
    clocks.stop()
    gpu_launches = eng.launches - launches0
    total_ms = sum(s.elapsed_time(e) for s, e in ev)

    # ---- timed region 2: host buffers through the public entry (e2e) ------------
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()              # synchronises the stream before returning
    barrier()
    e2e_s = time.perf_counter() - t0

    import frx
    total_ms = frx.sharding.max_over_ranks(total_ms, dev)      # the job is as slow as its slowest rank
    e2e_ms = frx.sharding.max_over_ranks(e2e_s * 1e3, dev)
    if rank != 0:
        return
    n_img = B * args.steps * world
    value = n_img / (total_ms / 1e3)
    e2e = n_img / (e2e_ms / 1e3)
    hbm, tflops, how = peaks()
    dec_s = statistics.mean(dec_ms) / 1e3
    enc_s = statistics.mean(enc_ms) / 1e3
    single_kernel = args.precision == "bf16" and statistics.mean(kern_ms) > 0
    if single_kernel:
        dec_s = statistics.mean(kern_ms) / 1e3
    dec_bytes = DECODE_BYTES_PER_IMAGE[args.precision] * B
    achieved = dec_bytes / dec_s / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16",
        "data": "synthetic",
        "config": workload_config(B, args.precision),
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": int(images_host.numel() * 4),
                "d2h_bytes_per_step": int(tokens_host.numel() * 8)},
        "gpu_launches": int(gpu_launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "traffic": ncu_traffic(DECODE_KERNEL, B) if single_kernel else None,
                     "kernel": DECODE_KERNEL + " (one launch = all 231 decode steps of the batch)" if single_kernel
                     else "greedy decode loop (CUDA graph of the fp32 step kernels)",
                     "algorithmic_bytes_per_launch": dec_bytes,
                     "ms": dec_s * 1e3, "peak_source": how},
        "roofline_encoder": {"bound": "tensor", "achieved": ENCODE_FLOP_PER_IMAGE * B / enc_s / 1e12, "peak": tflops,
                             "unit": "TFLOP/s", "frac": ENCODE_FLOP_PER_IMAGE * B / enc_s / 1e12 / tflops,
                             "ms": enc_s * 1e3, "peak_source": how},
        "clocks": clocks.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        # CPU leg (the only place bench.py touches oracle/): the oracle's op-for-op port of the reference's CPU
        # algorithm, timed on a bounded sample, on the BatchNorm-calibrated synthetic checkpoint of the parity
        # tests (plain random-init weights are numerically degenerate), and the token agreement of
        # both GPU modes with it on the same images.
        from oracle import satrn as o_satrn, synth as o_synth
        import frx
        from helpers import Vocab, flags_dict
        sd_cal = o_synth.synth_state_dict(o_satrn.ModelSpec(), 0)
        ips, cores, cpu_tok, dt = cpu_reference_throughput(sd_cal, args.cpu_sample, args.cpu_runs)
        xs = synthetic_images(args.cpu_sample, 0).to(dev)
        agree = {}
        with torch.no_grad():
            for prec in ("fp32", "bf16"):
                m = frx.EfficientSATRN(frx.Flags(flags_dict()).get(), Vocab(), sd_cal, None, precision=prec,
                                       max_batch=args.cpu_sample, max_steps=T).to(dev).eval()
                _, tok = m.greedy(xs, T)
                agree[prec] = (tok.cpu() == cpu_tok).float().mean().item()
                del m
        line["cpu_baseline"] = {
            "value": ips, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "%d batches of %d images x 231 decode steps (%.1f s), oracle port of the reference's "
                      "as-written CPU algorithm, torch %s fp32, BN-calibrated synthetic checkpoint"
                      % (args.cpu_runs, args.cpu_sample, dt, torch.__version__),
            "gpu_fp32_mode_token_agreement": agree["fp32"],
            "gpu_bf16_mode_token_agreement": agree["bf16"]}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="frx", choices=["frx", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--precision", default=os.environ.get("FRX_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--cpu-sample", type=int, default=32)
    ap.add_argument("--cpu-runs", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_frx(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
