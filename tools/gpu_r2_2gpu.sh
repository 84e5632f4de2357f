cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 900 python bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_2gpu_weak.json 2> /dev/null; echo "2 weak rc=$?"
timeout 600 python bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong --no-cpu-baseline --no-configs > gpurun_out/bench_2gpu_strong.json 2> /dev/null; echo "2 strong rc=$?"
python - <<'PY'
import json
for f in ["bench_2gpu_weak","bench_2gpu_strong"]:
    for l in open(f"gpurun_out/{f}.json"):
        if l.startswith("{"):
            d=json.loads(l); print(f, "%.4g"%d["value"], "%.4g"%d["e2e"]["value"], d.get("sharded_equals_single"))
PY
