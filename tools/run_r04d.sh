mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r04d_pytest_gpu.log 2>&1; tail -3 gpurun_out/r04d_pytest_gpu.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r04d_bench.json 2> gpurun_out/r04d_bench.err; tail -c 300 gpurun_out/r04d_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r04d_bench.json'))
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'ovw', d['e2e_overwrite']['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac'], d['checks']['all_ok'], d['clocks'], d['roofline']['timing'])
for s in d['sweep']: print(s['mesh'], s['states'], s.get('pairs_per_batch'), round(s['ms_per_step'],3), round(s['step_frac'],3))
print(d['psi_keep']['ms_per_step'])
P
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r04d_bench_reference.json 2>> gpurun_out/r04d_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r04d_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r04d_ncu_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
