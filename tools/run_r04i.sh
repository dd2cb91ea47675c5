mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r04i_pytest_gpu.log 2>&1; tail -2 gpurun_out/r04i_pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r04i_bench.json 2> gpurun_out/r04i_bench.err; tail -c 200 gpurun_out/r04i_bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r04i_bench_reference.json 2>> gpurun_out/r04i_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r04i_bench.json'))
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac'], d['checks']['all_ok'], d['clocks'])
print(d['cpu_baseline'])
r=json.load(open('gpurun_out/r04i_bench_reference.json')); print(r['value'], r['host_cores'], r['ms_per_step'])
P
