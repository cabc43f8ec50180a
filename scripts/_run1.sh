mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 6
for rep in 1 2; do
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(round(d['value'],1),round(d['e2e']['value'],1))"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'nms_|finalize|rank_decode|score_cand|select_topk' -s 138 -c 46 --csv --log-file gpurun_out/post_r2l.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
L=[l for l in open('gpurun_out/post_r2l.csv') if not l.startswith('==')]
d={}
for x in csv.DictReader(L):
    k=x['Kernel Name'][:22]; d.setdefault(k,[]).append(float(x['Metric Value'].replace(',',''))/1e3)
print({k:(round(sum(v),1), [round(t) for t in v[:6]]) for k,v in d.items()})
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nms_diag -s 54 -c 1 -o gpurun_out/r2l_diag python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
