cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
timeout 900 python bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_8gpu_weak.json 2> gpurun_out/r2_bench_8gpu_weak.err; echo "weak rc=$?"
timeout 600 python bench.py --gpus 8 --steps 5 --warmup 3 --scaling strong > gpurun_out/r2_bench_8gpu_strong.json 2> gpurun_out/r2_bench_8gpu_strong.err; echo "strong rc=$?"
timeout 900 python bench.py --gpus 8 --workload sweep --steps 2 --sweep-runs 100000 > gpurun_out/r2_sweep_8gpu.json 2> gpurun_out/r2_sweep_8gpu.err; echo "sweep rc=$?"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
tail -c 600 gpurun_out/r2_bench_8gpu_weak.err
