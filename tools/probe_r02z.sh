mkdir -p gpurun_out
{
for hb in 32 16 8 4; do echo "== CPB_HOST_BATCH=$hb"; CPB_HOST_BATCH=$hb python tools/e2e_probe.py 192 512 32 2>&1 | tail -4; done
} > gpurun_out/r02z_e2e_probe_batch.txt 2>&1
cat gpurun_out/r02z_e2e_probe_batch.txt
