mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 6
for rep in 1 2; do
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_bench_$rep.json 2> gpurun_out/r2k_bench.err; tail -c 300 gpurun_out/r2k_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r2k_bench_$rep.json'));print(d['value'],d['e2e']['value'],d['roofline']['all_convs']['forward_ms_sum_of_launches'],d['gpu_launches'])"
done
