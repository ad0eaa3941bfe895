set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -k "expert or defines or resize" 2>&1 | tail -12 > $O/s16_tests.log
timeout 300 ./imagemagick_b200/lib/shim_harness > $O/s16_shim.log 2>&1; echo "shim rc=$?" >> $O/s16_tests.log
cat $O/s16_tests.log; grep -E "FAIL|lobes|hits" $O/s16_shim.log | tail -6
