# usage: bash tools/run_variants.sh [-t] variant...   (-t: run the centred-scoring tests on each variant first)
T=0; if [ "$1" = "-t" ]; then T=1; shift; fi
for v in "$@"; do echo "== $v"; if [ "$v" = base ]; then unset RL_B200_LIB; else export RL_B200_LIB=recsys_lite_b200/librecsys_b200_$v.so; fi
  if [ $T = 1 ]; then timeout 300 python -m pytest tests/test_score_centred.py -x -q 2>&1 | tail -2; fi
  python tools/run_kernel.py score --users ${U:-303104} --items 100000 --mode 4 --iters ${ITERS:-20} --ir 1 --reps 3 2>&1 | tail -2; done
