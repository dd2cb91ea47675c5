# initcheck alone (after the fixes of the two oversized partial-sum copies), cases as in run_sanitizer.sh; inputs reach
# the device by memcpy because only the library's kernels are instrumented (a tensor written by a torch kernel would
# read as uninitialised)
mkdir -p gpurun_out
OUT=gpurun_out/r04g_initcheck.txt
: > $OUT
for c in "24 6 full" "48 5 full" "192 2"; do
  echo "== initcheck: $c" >> $OUT
  timeout 200 compute-sanitizer --tool initcheck --kernel-regex kns=cpb --print-limit 50 python tools/sanitize_case.py $c > gpurun_out/r04g_initcheck_raw.txt 2>&1
  grep -E "sanitize case ok|ERROR SUMMARY|AssertionError" gpurun_out/r04g_initcheck_raw.txt >> $OUT
  grep -E "^========= (Uninitialized|Host API)|Device Frame.*kernels|Host Frame: .*(cpb_|_impl)" gpurun_out/r04g_initcheck_raw.txt | sed 's/0x[0-9a-f]*/ADDR/g' | sort | uniq -c | sort -rn | head -8 >> $OUT
done
cat $OUT
