mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:k_zw --launch-skip 2 --launch-count 2 -o gpurun_out/prof_r02k_zw -f python tools/gpu_probe.py 192 128 32 1 > gpurun_out/prof_r02k_zw.log 2>&1
tail -5 gpurun_out/prof_r02k_zw.log
