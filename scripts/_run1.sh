mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'conv_tc_kernel<\(int\)256, \(int\)2, \(int\)1>' -s 47 -c 1 -o gpurun_out/r2j_res4conv3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_ncu_res4.log 2>&1
tail -n 2 gpurun_out/r2j_ncu_res4.log | cut -c1-200
