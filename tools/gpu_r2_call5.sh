cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12
python tools/probe_multi.py 100000:10000:10:101:0 100000:1250:10:101:0 100000:1300:10:101:0 100000:2368:10:101:0 10000:10000:10:100:0 10000:2500:10:100:0
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json
timeout 600 python bench.py --workload sweep --steps 1 --sweep-runs 8000 > gpurun_out/sweep_quick.json 2> gpurun_out/sweep_quick.err; echo "sweep rc=$?"; tail -3 gpurun_out/sweep_quick.err; cat gpurun_out/sweep_quick.json
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/ref_quick.json 2> gpurun_out/ref_quick.err; echo "ref rc=$?"; tail -3 gpurun_out/ref_quick.err; cat gpurun_out/ref_quick.json
