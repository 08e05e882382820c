# ncu evidence of the final r2 kernels (one GPU).  Each capture only after the same command exited 0 without ncu.
set -x
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-api > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2b_launches_ncu.csv \
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-api > gpurun_out/ncu_bench.log 2>&1
python tools/run_kernel.py score --users 1000000 --items 100000 --mode 4 --iters 25 --ir 1 --reps 1 > gpurun_out/plain_rk.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"score_tc2_kernel|als_k64_kernel" -s 48 -c 3 -f -o gpurun_out/r2b_full \
    python tools/run_kernel.py score --users 1000000 --items 100000 --mode 4 --iters 25 --ir 1 --reps 1 > gpurun_out/ncu_rk.log 2>&1
tail -2 gpurun_out/plain_rk.log; ls -la gpurun_out/r2b_*
