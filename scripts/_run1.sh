mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_dense_gpu.py tests/test_detector_gpu.py -x -q -m gpu 2>&1 | tail -n 4
run() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-json gpurun_out/ab_$tag.json > gpurun_out/ab_bench_$tag.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/ab_bench_$tag.json'));p=json.load(open('gpurun_out/ab_$tag.json'))
def S(f): return round(sum(o['ms'] for o in p if f(o['name']))*1e3,1)
print('$tag', round(d['value'],1), round(d['e2e']['value'],1), 'fwd',round(d['roofline']['all_convs']['forward_ms_sum_of_launches'],3),'res2c3',S(lambda n:'res2' in n and n.endswith('conv3')),'res3c3',S(lambda n:'res3' in n and n.endswith('conv3')),'res4c3',S(lambda n:'res4' in n and n.endswith('conv3')),'res5c3',S(lambda n:'res5' in n and n.endswith('conv3')),'lat',S(lambda n:'lateral' in n))"
}
run policy A=1
run old DAFNE_CONV_RES_BN=256 DAFNE_CONV_RING=3
run ring4 DAFNE_CONV_RING=4
run ring5 DAFNE_CONV_RING=5
run policy2 A=1
run ring3 DAFNE_CONV_RING=3
