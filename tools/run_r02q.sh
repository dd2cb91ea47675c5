mkdir -p gpurun_out
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
tail -c 600 gpurun_out/r02q_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r02q_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac'], d['checks']['all_ok'])
print(d['roofline']['kernel_ms_per_step'])
for s in d['sweep']: print(s['mesh'], s['states'], s.get('pairs_per_batch'), round(s['ms_per_step'],3), round(s['step_frac'],3))
P
