# compute-sanitizer over small cases of the hot path (memcheck / racecheck / synccheck / initcheck).
# Kernel coverage: 24 = 6 x 4 and 48 = 12 x 4 radices (staged TMA y/z kernels, mirror-pair x kernels with cp.async),
# device- and host-pointer entry points, LSD.  Only the library's own kernels are instrumented (--kernel-regex).
mkdir -p gpurun_out
OUT=gpurun_out/r04f_sanitizer.txt
: > $OUT
KR="--kernel-regex kns=cpb"
for tool in memcheck synccheck initcheck; do
  for c in "24 6 full" "48 5 full" "192 2"; do
    echo "== $tool: $c" >> $OUT
    timeout 240 compute-sanitizer --tool $tool $KR --print-limit 5 python tools/sanitize_case.py $c 2>&1 | grep -E "sanitize case ok|ERROR SUMMARY|Error|error|hazard|=========     at|AssertionError" | head -12 >> $OUT
  done
done
for c in "24 6" "48 5"; do
  echo "== racecheck: $c" >> $OUT
  timeout 280 compute-sanitizer --tool racecheck $KR --racecheck-report all --print-limit 5 python tools/sanitize_case.py $c 2>&1 | grep -E "sanitize case ok|RACECHECK SUMMARY|ERROR SUMMARY|hazard|Error|=========     at|AssertionError" | head -12 >> $OUT
done
cat $OUT
