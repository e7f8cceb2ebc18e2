#!/bin/bash
# Profiling pass of one round (run on the GPU box through gpurun):  bash scripts/profile_round.sh r02
# 1) launch list of the default bench command (shares of the step + DRAM bytes per launch), 2) ncu --set full of the tensor-core
# layer kernel at ranks 256 / 128 / 32, 3) of the batch-1 wavefront kernel, 4) of the units = 1024 kernel (C5 shard) + the C5
# launch list.  Every ncu run is preceded by the identical plain command (B200_PROFILING.md rule).  Summaries for profiles/ are
# produced afterwards on the CPU box by scripts/make_profile_summaries.py.
R=${1:-r02}
O=gpurun_out
# gpurun brings back at most 64 MiB: condense every report ON the box (counters + the 40 hottest source lines by stall samples) and
# keep only the rank-128 report itself
summarise() {
  python scripts/ncu_summary.py $O/$1.ncu-rep $2 > $O/$1.txt 2>/dev/null
  ncu -i $O/$1.ncu-rep --page source --csv 2>/dev/null | python scripts/ncu_hot_lines.py > $O/$1.hot.txt 2>/dev/null
  case $1 in tc_layer_rank128_*) ;; *) rm -f $O/$1.ncu-rep ;; esac
}
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep --no-batch1 --no-c4"
$BENCH > $O/plain_bench_$R.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file $O/launches_$R.csv $BENCH > $O/ncu_bench_$R.log 2>&1
for rk in 256 128 32; do
  CMD="python scripts/tc_time.py $rk 4096 128"
  $CMD > $O/plain_tc_$rk.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:lstm_tc_ -s 2 -c 1 -f -o $O/tc_layer_rank${rk}_$R $CMD > $O/ncu_tc_$rk.log 2>&1
  summarise tc_layer_rank${rk}_$R lstm_tc
done
CMD="python scripts/prof_batch1.py"
$CMD > $O/plain_b1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lstm_wavefront_kernel -s 1 -c 1 -f -o $O/b1_wavefront_$R $CMD > $O/ncu_b1.log 2>&1
summarise b1_wavefront_$R lstm_wavefront
# the FP32 (parity) engine at the headline batch
CMD="python scripts/fp32_engine_time.py"
$CMD > $O/plain_fp32.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lstm_general_kernel -s 1 -c 1 -f -o $O/general_fp32_$R $CMD > $O/ncu_fp32.log 2>&1
summarise general_fp32_$R lstm_general
for f in $O/plain_tc_128.log $O/plain_tc_32.log $O/plain_b1.log; do tail -n 2 $f; done
# C5 shard: launch list (K2 Jacobi, K3 fused penalties, packing, the units = 1024 tensor-core kernel) at T = 128, then the full set on that kernel
CMD="python scripts/c5_parts.py 128"
$CMD > $O/plain_c5_$R.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'penalties_kernel|lstm_tc_|pack_' -c 60 --csv --log-file $O/launches_c5_$R.csv $CMD > $O/ncu_c5_$R.log 2>&1
CMD="python scripts/c5_parts.py 64"
$CMD > $O/plain_c5b_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lstm_tc_pipe_kernel -s 1 -c 1 -f -o $O/tc_pipe_c5_$R $CMD > $O/ncu_c5b_$R.log 2>&1
summarise tc_pipe_c5_$R lstm_tc
tail -n 1 $O/plain_c5_$R.log | cut -c1-300
