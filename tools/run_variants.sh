# usage: bash tools/run_variants.sh [-t] variant...   (-t: run the centred-scoring tests on each variant first)
T=0; if [ "$1" = "-t" ]; then T=1; shift; fi
for v in "$@"; do echo "== $v"; export RL_B200_LIB=recsys_lite_b200/librecsys_b200_$v.so
  if [ $T = 1 ]; then timeout 300 python -m pytest tests/test_score_centred.py -x -q 2>&1 | tail -2; fi
  python tools/run_kernel.py score --users 302000 --items 100000 --mode 4 --iters 12 --reps 3 2>&1 | tail -2; done
