timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2o.err | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(round(d['value'],1),round(d['e2e']['value'],1),d['clocks'])"
tail -n 3 gpurun_out/r2o.err
