cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2_bench_final.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print("value %.4g e2e %.4g frac %.3g cpu %.3g clocks %s sharded %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["clocks"], d.get("sharded_equals_single")))
        for c in d.get("configs", []): print("  ", {k: c[k] for k in c if k in ("name","value","config","events_per_s")})
PY
