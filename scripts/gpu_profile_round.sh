#!/bin/bash
# One gpurun call that refreshes the tracked evidence under profiles/ (run from the repo root ON THE GPU BOX):
#   scripts/gpu_profile_round.sh <tag>      e.g. r3k
# Writes gpurun_out/<tag>_*; scripts/summarize_profiles.py turns those into profiles/<tag>_*.md here afterwards.
# The default workload of bench.py is r101_b32 (BASELINE.json configs[2]): 115 tcgen05 launches per forward (conv_tc, stem_tc,
# tail_tc, pair_tc).
set -u
TAG=${1:-rX}
mkdir -p gpurun_out
# 1. bench line + per-launch events (the numbers; never taken under a profiler)
timeout 500 python bench.py --profile-json gpurun_out/${TAG}_per_launch_events_r101_b32.json > gpurun_out/${TAG}_bench_r101_b32.json 2> gpurun_out/${TAG}_bench.err
# 1b. the alternates (profiles/ only)
for w in r50_b8 hrsc_r50_mixed r101_b32_nmsmax; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${w}.json 2>> gpurun_out/${TAG}_bench.err
done
# 2. launch list of the same command (shares)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${TAG}_launches_r101_b32.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
# 3. tensor-pipe utilisation of every conv launch of one forward (forward 4 of the run: skip the 3 warm-up forwards);
#    the number of tcgen05 launches per forward comes from step 1's per-launch profile
NCONV=$(python -c "import json; print(sum(1 for o in json.load(open('gpurun_out/${TAG}_per_launch_events_r101_b32.json')) if o['kind'] in (1, 2)))")
echo "tcgen05 launches per forward: $NCONV"
timeout 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:'conv_tc|stem_tc|tail_tc|pair_tc' -s $((3 * NCONV)) -c $NCONV --csv --log-file gpurun_out/${TAG}_conv_tensor_pipe_r101_b32.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_pipe.log 2>&1
# 4. full captures: the dominant kernel (tower conv, second layer: GroupNorm on load + GN-statistics epilogue), the NMS
#    broadcast (panel 0, second pass) and the CTA-pair kernel (conv1 of a res4 block)
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'conv_tc_kernel<\(int\)256, \(int\)1, \(int\)2>' -s 25 -c 1 -o gpurun_out/${TAG}_tower \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_tower.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nms_bcast -s 103 -c 1 -o gpurun_out/${TAG}_nms_bcast \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_nms.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pair_tc -s 90 -c 1 -o gpurun_out/${TAG}_pair_conv1 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_pair.log 2>&1
cut -c1-300 gpurun_out/${TAG}_bench_r101_b32.json
tail -n 2 gpurun_out/${TAG}_ncu_tower.log gpurun_out/${TAG}_ncu_pipe.log gpurun_out/${TAG}_ncu_nms.log
