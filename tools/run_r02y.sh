mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02y_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02y_pytest_gpu.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02y_bench.json 2> gpurun_out/r02y_bench.err; tail -c 300 gpurun_out/r02y_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02y_bench_reference.json 2>> gpurun_out/r02y_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02y_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02y_ncu_bench.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:"k_z_vpsi|k_z_rho" --launch-skip 4 --launch-count 4 -o gpurun_out/prof_r02y_z -f python tools/gpu_probe.py 192 128 32 1 > gpurun_out/prof_r02y_z.log 2>&1
python - <<'P'
import json
d=json.load(open('gpurun_out/r02y_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac'], d['checks']['all_ok'], d['clocks'])
print(d['roofline']['kernel_ms_per_step'])
for s in d['sweep']: print(s['mesh'], s['states'], s.get('pairs_per_batch'), round(s['ms_per_step'],3), round(s['step_frac'],3))
print(open('gpurun_out/r02y_bench_reference.json').read()[:600])
P
