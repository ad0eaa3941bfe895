set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -k "xyz_family or r02b" 2>&1 | tail -25 > $O/s15_tests.log
cat $O/s15_tests.log
