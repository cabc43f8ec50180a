mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 4
bash scripts/gpu_profile_round.sh r2n
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload r101_b32 > gpurun_out/r2n_bench_r101_b32.json 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/r2n_bench_r101_b32.json'));print('r101_b32',round(d['value'],1),round(d['e2e']['value'],1))"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
