mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r03g_bench_${N}gpu.json 2> gpurun_out/r03g_bench_${N}gpu.err
tail -c 300 gpurun_out/r03g_bench_${N}gpu.err
python - <<P
import json
d=json.load(open('gpurun_out/r03g_bench_${N}gpu.json'))
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['checks']['all_ok'], d['checks']['ranks_bit_identical'], d['clocks'])
k=d['roofline']['kernel_ms_per_step']; print(sum(k.values()), k)
P
