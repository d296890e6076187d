#!/bin/bash
# Everything that needs more than one GPU, in one gpurun call (N GPUs of one box; results under gpurun_out/):
#   bash acg-alp-ldpc_b200/tools/multi_gpu_session.sh 8
N=${1:-8}
FRAMES3=${FRAMES3:-1000000000}
OPT_ITERS=${OPT_ITERS:-1000}
FULL=${FULL:-1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/r02_topo_n$N.txt 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu_entry" > $OUT/r02_multi_entry_n$N.log 2>&1; tail -1 $OUT/r02_multi_entry_n$N.log
run_bench() {   # ranks, extra flags, tag
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $1 --steps 5 --warmup 3 $2 > $OUT/r02_bench_$3.json 2> $OUT/r02_bench_$3.err
  echo "bench $3 rc=$?"; python acg-alp-ldpc_b200/tools/show_bench.py $OUT/r02_bench_$3.json | head -3
}
if [ $FULL = 1 ]; then run_bench $N "" n$N; else run_bench $N "--headline-only --with-experiment --no-cpu-baseline" n$N; fi
for k in 4 2; do if [ $k -lt $N ]; then run_bench $k "--headline-only --with-experiment --no-cpu-baseline" n${k}_of$N; fi; done
python bench.py --gpus 1 --steps 5 --warmup 3 --headline-only --with-experiment --no-cpu-baseline > $OUT/r02_bench_n1_of$N.json 2> $OUT/r02_bench_n1_of$N.err; python acg-alp-ldpc_b200/tools/show_bench.py $OUT/r02_bench_n1_of$N.json | head -1
python acg-alp-ldpc_b200/tools/config3_run.py --frames $FRAMES3 > $OUT/r02_config3_1e9_n$N.txt 2> $OUT/r02_config3_1e9_n$N.err; cat $OUT/r02_config3_1e9_n$N.txt | cut -c1-160
cd acg-alp-ldpc_b200 && g++ -std=c++17 -pthread -O2 -I. -I../include optimize_H.cpp -o /tmp/optimize_H -L. -lldpc_b200 -Wl,-rpath,$PWD
for g in 1 $N; do
  T0=$(date +%s%N)
  LDPC_GPUS=$g LDPC_OPT_ITERS=$OPT_ITERS LDPC_OPT_SAVE=/tmp/opt_g$g.txt LDPC_OPT_START=data/H05 /tmp/optimize_H > ../$OUT/r02_optimize_H_${OPT_ITERS}_gpus$g.txt 2> ../$OUT/r02_optimize_H_${OPT_ITERS}_gpus$g.err
  echo "optimize_H LDPC_OPT_ITERS=$OPT_ITERS gpus=$g rc=$? wall $(( ($(date +%s%N) - T0) / 1000000 )) ms (includes the final 10000-frame FER)" | tee -a ../$OUT/r02_optimize_H_timing_n$N.txt
done
cmp ../$OUT/r02_optimize_H_${OPT_ITERS}_gpus1.txt ../$OUT/r02_optimize_H_${OPT_ITERS}_gpus$N.txt && cmp /tmp/opt_g1.txt /tmp/opt_g$N.txt && echo "optimize_H: identical trajectory and matrix on 1 and $N GPUs" | tee -a ../$OUT/r02_optimize_H_timing_n$N.txt
