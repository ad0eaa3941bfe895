set -u
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 > $O/s11_tests.log
timeout 300 ./imagemagick_b200/lib/shim_harness > $O/s11_shim.log 2>&1; echo "shim rc=$?" >> $O/s11_tests.log
cat $O/s11_tests.log; grep -E "FAIL|hits" $O/s11_shim.log | tail -6
