mkdir -p gpurun_out
{
for m in 0 0x7f 0x07 0x08 0x10 0x20 0x01 0x37; do echo "== CPB_PDL=$m"; CPB_PDL=$m timeout 300 python tools/gpu_probe.py 192 256 32 4 2>&1 | tail -4 | head -2; done
} > gpurun_out/r02p_probe_pdl_mask.txt 2>&1
cat gpurun_out/r02p_probe_pdl_mask.txt
