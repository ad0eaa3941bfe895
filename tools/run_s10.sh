set -u
O=gpurun_out
timeout 300 python tools/diag_xyzfamily.py > $O/s10_diag.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -k "xyz_family or r02b" 2>&1 | tail -12 > $O/s10_tests.log
cat $O/s10_diag.log; cat $O/s10_tests.log
