cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 900 python bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2_bench_4gpu_weak.json 2> gpurun_out/r2_bench_4gpu_weak.err; echo "weak rc=$?"
timeout 600 python bench.py --gpus 4 --steps 5 --warmup 3 --scaling strong > gpurun_out/r2_bench_4gpu_strong.json 2> gpurun_out/r2_bench_4gpu_strong.err; echo "strong rc=$?"
timeout 900 python bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_weak.json 2> /dev/null; echo "2 weak rc=$?"
timeout 600 python bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong > gpurun_out/bench_2gpu_strong.json 2> /dev/null; echo "2 strong rc=$?"
for f in r2_bench_4gpu_weak r2_bench_4gpu_strong bench_2gpu_weak bench_2gpu_strong; do grep "^{" gpurun_out/$f.json | cut -c1-230; done
