# multi-GPU lines on one 8-GPU box: C4 at N = 8, 4, 2 (driver-style); pass C5=1 for the C5 line at N = 8 as well
set -x
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@"; }
run 8 --steps 20 --warmup 5 > gpurun_out/b_n8.json 2> gpurun_out/b_n8.err
run 4 --steps 20 --warmup 5 --no-cpu > gpurun_out/b_n4.json 2> gpurun_out/b_n4.err
if [ -n "$N2" ]; then run 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/b_n2.json 2> gpurun_out/b_n2.err; fi
if [ -n "$C5" ]; then run 8 --workload C5 --steps 5 --warmup 3 > gpurun_out/bench_c5_n8.json 2> gpurun_out/bench_c5_n8.err; fi
python - <<'P'
import json
for f in ("b_n8","b_n4","b_n2","bench_c5_n8"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d["ms"], (d.get("e2e") or {}).get("ms_per_step"), (d.get("parity") or {}).get("ok"))
    except Exception as e: print(f, "failed", e)
P
