set -x
P=acg-alp-ldpc_b200/tools/profile_case.py
run() {  # tag, kernel regex, args...
  tag=$1; rx=$2; shift 2
  python $P "$@" > gpurun_out/plain_$tag.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o gpurun_out/prof_r2_$tag -f python $P "$@" > gpurun_out/ncu_r2_$tag.log 2>&1
  tail -1 gpurun_out/plain_$tag.log
  python acg-alp-ldpc_b200/tools/ncu_summary.py gpurun_out/prof_r2_$tag.ncu-rep $UNITS > gpurun_out/r02_${tag}_ncu.txt 2>&1
  if [ "$KEEP" != 1 ]; then rm -f gpurun_out/prof_r2_$tag.ncu-rep; fi
}
UNITS=$((23680*100*860)) KEEP=1; run bp_H05 bp_lr --algo bp --code H05 --frames 23680
UNITS=$((5920*100*3024)) KEEP=0; run bp_1008 bp_lr --algo bp --code reg_3_6_1008 --frames 5920
UNITS=$((4736*200*580)) KEEP=1; run admm_optimalH qpadmm_chk --algo qpadmm --code optimalH --frames 4736 --iters 200
UNITS=$((1184*200*2016)) KEEP=0; run admm_1008 qpadmm_chk --algo qpadmm --code reg_3_6_1008 --frames 1184 --iters 200
python bench.py --steps 2 --warmup 1 --frames 131072 --exp-frames 262144 --no-cpu-baseline > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --frames 131072 --exp-frames 262144 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/ | tail -20; du -sh gpurun_out
