mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "async or device_entry_points_match or production_shape" > gpurun_out/r02s_pytest_async.log 2>&1; tail -3 gpurun_out/r02s_pytest_async.log
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
tail -c 300 gpurun_out/r02s_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r02s_bench.json'))
print(d['ms_per_step'], d['value'], d['checks']['all_ok'], d['clocks'], d['roofline']['ms_per_step_serialised'])
P
