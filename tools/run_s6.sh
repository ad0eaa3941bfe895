set -u
O=gpurun_out
for n in 2 3; do
timeout 600 python bench.py --streams $n > $O/s6_bench_$n.json 2> $O/s6_bench_$n.err
done
python - <<'PY'
import json
for n in (2,3):
    f=f"gpurun_out/s6_bench_{n}.json"
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d["config"]
        print(n, d["value"], d["ms_per_step"], c["timing"], c["blur_ms"], c["resize_ms"], d.get("clocks"))
    except Exception as e:
        print(n, "failed", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
