# end-state multi-GPU confirmation: peer-collective tests and the bench line with its checks on N GPUs of one box
mkdir -p gpurun_out
N=${1:-8}
timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r04e_pytest_multi_${N}gpu.log 2>&1; tail -3 gpurun_out/r04e_pytest_multi_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r04e_bench_${N}gpu.json 2> gpurun_out/r04e_bench_${N}gpu.err
tail -c 300 gpurun_out/r04e_bench_${N}gpu.err
python - <<P
import json
d=json.load(open('gpurun_out/r04e_bench_${N}gpu.json'))
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['checks']['all_ok'], d['checks']['ranks_bit_identical'], d['clocks'])
P
