set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hit_and_miss or errors_are_loud or morph or dilate or erode" 2>&1 | tail -12 > $O/s7_tests.log
timeout 300 ./imagemagick_b200/lib/shim_harness > $O/s7_shim.log 2>&1; echo "shim rc=$?" >> $O/s7_tests.log
cat $O/s7_tests.log; grep -E "FAIL|HSL|HSV|Jinc|Kaiser|HitAndMiss|Thinning|Intensity|Distance|hits" $O/s7_shim.log | tail -14
