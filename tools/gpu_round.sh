# One GPU-box pass: smoke, GPU tests, bench (both arms).  Usage: bash tools/gpu_round.sh
cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
nproc
python __graft_entry__.py --smoke 2>&1 | tail -3
python -m pytest tests -m gpu -q 2>&1 | tail -5
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
