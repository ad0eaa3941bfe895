set -u
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/s2_tests.log
timeout 600 python bench.py > $O/s2_bench.json 2> $O/s2_bench.err; echo "bench rc=$?" >> $O/s2_tests.log
cat $O/s2_tests.log; head -c 1500 $O/s2_bench.json
