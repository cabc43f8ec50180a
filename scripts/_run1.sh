mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 4
bash scripts/gpu_profile_round.sh r2m
for w in r101_b32 hrsc_r50_b8 hrsc_r50_mixed hrsc_r50_bucketed; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/r2m_bench_$w.json 2> gpurun_out/r2m_bench_$w.err; tail -c 300 gpurun_out/r2m_bench_$w.err
python -c "
import json;d=json.load(open('gpurun_out/r2m_bench_$w.json'));print('$w',round(d['value'],1),round(d['e2e']['value'],1))"
done
