#!/bin/bash
# One gpurun call that refreshes the tracked evidence under profiles/ (run from the repo root ON THE GPU BOX):
#   scripts/gpu_profile_round.sh <tag>      e.g. r1k
# Writes gpurun_out/<tag>_*; scripts/summarize_profiles.py turns those into profiles/<tag>_*.md here afterwards.
set -u
TAG=${1:-rX}
mkdir -p gpurun_out
# 1. bench line + per-launch events (the numbers; never taken under a profiler)
timeout 400 python bench.py --profile-json gpurun_out/${TAG}_per_launch.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
# 2. launch list of the same command (shares)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
# 3. tensor-pipe utilisation of every conv launch of one forward (step 4 of 4: skip the 3 warm-up forwards)
timeout 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:'conv_tc|stem_tc' -s 207 -c 69 --csv --log-file gpurun_out/${TAG}_conv_tensor_pipe.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_pipe.log 2>&1
# 4. full captures: the dominant kernel (tower conv, GN-statistics epilogue) and the NMS broadcast of panel 0
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'conv_tc_kernel<\(int\)256, \(int\)1, \(int\)2>' -s 24 -c 1 -o gpurun_out/${TAG}_tower \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_tower.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nms_bcast -s 51 -c 1 -o gpurun_out/${TAG}_nms_bcast \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_nms.log 2>&1
cut -c1-300 gpurun_out/${TAG}_bench.json
tail -n 2 gpurun_out/${TAG}_ncu_tower.log gpurun_out/${TAG}_ncu_pipe.log
