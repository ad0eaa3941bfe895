set -u
O=gpurun_out
tools/micro/dmma_feed > $O/dmma_feed.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/s4_tests.log
cat $O/dmma_feed.log $O/s4_tests.log
