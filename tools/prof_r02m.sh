mkdir -p gpurun_out
CPB_XW=3 timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:k_xw --launch-skip 4 --launch-count 3 -o gpurun_out/prof_r02m_xw -f python tools/gpu_probe.py 192 128 32 1 > gpurun_out/prof_r02m_xw.log 2>&1
tail -3 gpurun_out/prof_r02m_xw.log
