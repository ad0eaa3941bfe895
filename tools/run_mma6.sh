set -u
O=gpurun_out
timeout 600 python bench.py > $O/s3_bench_auto.json 2> $O/s3_bench_auto.err
MB200_MMA=1 timeout 600 python bench.py > $O/s3_bench_mma.json 2> $O/s3_bench_mma.err
python - <<'PY'
import json
for f in ("gpurun_out/s3_bench_auto.json","gpurun_out/s3_bench_mma.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    c=d["config"]
    print(f, d["value"], d["ms_per_step"], c["blur_ms"], c["resize_ms"], d.get("clocks"), d["roofline"])
PY
