cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 600 python bench.py --gpus 8 --steps 5 --warmup 3 --scaling strong > gpurun_out/r2_bench_8gpu_strong.json 2> gpurun_out/r2_bench_8gpu_strong.err; echo "strong rc=$?"
grep "^{" gpurun_out/r2_bench_8gpu_strong.json | cut -c1-260
