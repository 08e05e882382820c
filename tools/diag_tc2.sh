# scoring diagnostics on converged factors: plain timing, debug counters, ceilings (RL_TC2_FLAGS)
set -x
U=${U:-303104}
python -m pytest tests/test_score_centred.py -x -q -m gpu > gpurun_out/d_test.log 2>&1; tail -3 gpurun_out/d_test.log
python tools/run_kernel.py score --users $U --items 100000 --mode 4 --iters 20 --reps 2 --ir 1 --every 5 > gpurun_out/d_plain.log 2>&1
for f in ${FLAGS:-0}; do
RL_TC_DEBUG=1 RL_TC2_FLAGS=$f python tools/run_kernel.py score --users $U --items 100000 --mode 4 --iters 20 --reps 2 --ir 1 > gpurun_out/d_flag$f.log 2>&1
done
grep -v "^+" gpurun_out/d_plain.log | tail -8
grep "sel warp 0\|scan warp 0\|mma issuer 0\|rep 1" gpurun_out/d_flag0.log | tail -4
