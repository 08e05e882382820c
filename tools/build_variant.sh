#!/bin/bash
# tools/build_variant.sh NAME "-DT2_X=1 ..." : librecsys_b200_NAME.so with score_tc2.cu compiled under the given defines
# (kernel-variant experiments; run with RL_B200_LIB=recsys_lite_b200/librecsys_b200_NAME.so)
set -e
cd "$(dirname "$0")/.."
NAME=$1; DEFS=$2
python -m recsys_lite_b200.build > /dev/null
B=recsys_lite_b200/build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr $DEFS \
  -c recsys_lite_b200/csrc/score_tc2.cu -o $B/score_tc2_$NAME.o
OBJS=$(ls $B/*.o | grep -v "score_tc2")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o recsys_lite_b200/librecsys_b200_$NAME.so $OBJS $B/score_tc2_$NAME.o -Xcompiler -fPIC
echo recsys_lite_b200/librecsys_b200_$NAME.so
