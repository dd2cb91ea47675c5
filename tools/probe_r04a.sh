# does an L2-sized batch make the HBM-bound y passes faster? (T1 of 1 / 2 / 4 pairs = 22 / 44 / 88 MB against 126 MB of L2)
mkdir -p gpurun_out
{
for mb in 1 2 4 8 32; do echo "== 192^3 x 128 states, batch $mb, 1 stream"; CPB_STREAMS=1 timeout 300 python tools/gpu_probe.py 192 128 $mb 2 2>&1 | tail -3; done
for mb in 2 4; do echo "== 192^3 x 128 states, batch $mb, 2 streams"; timeout 300 python tools/gpu_probe.py 192 128 $mb 2 2>&1 | tail -3; done
} > gpurun_out/r04a_probe_l2batch.txt 2>&1
cat gpurun_out/r04a_probe_l2batch.txt
