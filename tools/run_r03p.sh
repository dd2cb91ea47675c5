mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r03p_pytest_multi_${N}gpu.log 2>&1; tail -3 gpurun_out/r03p_pytest_multi_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r03p_bench_${N}gpu.json 2> gpurun_out/r03p_bench_${N}gpu.err
tail -c 400 gpurun_out/r03p_bench_${N}gpu.err
CPB_BCAST_OVERLAP=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r03p_bench_${N}gpu_nooverlap.json 2>> gpurun_out/r03p_bench_${N}gpu.err
python - <<P
import json
for f in ('gpurun_out/r03p_bench_${N}gpu.json','gpurun_out/r03p_bench_${N}gpu_nooverlap.json'):
    d=json.load(open(f))
    print(f, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['checks']['all_ok'], d['checks']['ranks_bit_identical'], d['clocks'])
P
