cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
grep "^{" gpurun_out/r2_bench_final.json | cut -c1-400
