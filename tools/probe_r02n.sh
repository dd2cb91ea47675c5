mkdir -p gpurun_out
{
for mb in 32 48 64; do echo "== 192^3 x 256 states, batch $mb"; timeout 300 python tools/gpu_probe.py 192 256 $mb 2 2>&1 | tail -3; done
for mb in 32 64; do echo "== 96^3 x 128 states, batch $mb"; timeout 300 python tools/gpu_probe.py 96 128 $mb 3 2>&1 | tail -3; done
for mb in 32 64; do echo "== 120^3 x 128 states, batch $mb"; timeout 300 python tools/gpu_probe.py 120 128 $mb 3 2>&1 | tail -3; done
for mb in 8 16; do echo "== 320^3 x 32 states, batch $mb"; timeout 300 python tools/gpu_probe.py 320 32 $mb 2 2>&1 | tail -3; done
} > gpurun_out/r02n_probe_batch.txt 2>&1
cat gpurun_out/r02n_probe_batch.txt
