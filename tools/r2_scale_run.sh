# 8-GPU lines of every workload (run with gpurun --gpus 8); one rank per GPU under torchrun, NCCL
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_greedy_n8.json 2> gpurun_out/r2_bench_greedy_n8.err
$TR bench.py --gpus 8 --workload train --steps 10 --warmup 3 > gpurun_out/r2_bench_train_n8.json 2> gpurun_out/r2_bench_train_n8.err
$TR bench.py --gpus 8 --workload lite --steps 10 --warmup 3 > gpurun_out/r2_bench_lite_n8.json 2> gpurun_out/r2_bench_lite_n8.err
$TR bench.py --gpus 8 --workload swin --steps 5 --warmup 3 > gpurun_out/r2_bench_swin_n8.json 2> gpurun_out/r2_bench_swin_n8.err
$TR bench.py --gpus 8 --workload beam4 --steps 5 --warmup 3 > gpurun_out/r2_bench_beam4_n8.json 2> gpurun_out/r2_bench_beam4_n8.err
