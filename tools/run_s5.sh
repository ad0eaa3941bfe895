set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu -k "hexcone or r02b or resize_all_filters or errors_are_loud" 2>&1 | tail -12 > $O/s5_tests.log
timeout 300 ./imagemagick_b200/lib/shim_harness > $O/s5_shim.log 2>&1; echo "shim rc=$?" >> $O/s5_tests.log
cat $O/s5_tests.log; grep -E "FAIL|Colorspace|Jinc|Kaiser|hits|fallback" $O/s5_shim.log | tail -12
