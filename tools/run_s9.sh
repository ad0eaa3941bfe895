set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu -k "xyz_family or r02b or hexcone or colorspace" 2>&1 | tail -15 > $O/s9_tests.log
timeout 300 ./imagemagick_b200/lib/shim_harness > $O/s9_shim.log 2>&1; echo "shim rc=$?" >> $O/s9_tests.log
cat $O/s9_tests.log; grep -E "FAIL|Luv|hits" $O/s9_shim.log | tail -6
