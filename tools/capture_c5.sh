# ncu evidence of the C5 user half-step (dual-form kernel) and item half-step (primal k = 128 kernel): one rank's shard on one GPU
set -x
export RL_BENCH_ALLOW_SHORT_WARMUP=1
python bench.py --workload C5 --emulate-world 8 --steps 1 --warmup 1 > gpurun_out/plain_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"als_k128" -s 4 -c 1 -f -o gpurun_out/r2_c5u \
    python bench.py --workload C5 --emulate-world 8 --steps 1 --warmup 1 > gpurun_out/ncu_c5.log 2>&1
ls -la gpurun_out/r2_c5u.ncu-rep
