set -u
O=gpurun_out
timeout 80 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_r02.py tests/test_gpu_parity.py -q -m gpu -x \
  -k "rank_statistics_on_few_levels and (win1 or win2) and not 4-1 or colorspace_settings and noise" > $O/r02c_sanitizer.log 2>&1
echo "rc=$?" >> $O/r02c_sanitizer.log
tail -5 $O/r02c_sanitizer.log
