cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 1200 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench.err
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"
