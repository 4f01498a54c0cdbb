set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:dec_cluster -c 1 -o gpurun_out/r2_dec_p2 -f python tools/profile_pass.py --precision bf16 > gpurun_out/r2_ncu_dec.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"tc_igemm_ws|dwconv4_se_mean" -c 24 -o gpurun_out/r2_enc -f python tools/profile_pass.py --precision bf16 --steps 2 > gpurun_out/r2_ncu_enc.log 2>&1
ls -la gpurun_out/*.ncu-rep
