cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -8
timeout 600 python bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_weak.json 2> gpurun_out/bench_2gpu_weak.err; echo "weak rc=$?"; tail -2 gpurun_out/bench_2gpu_weak.err; cat gpurun_out/bench_2gpu_weak.json
timeout 600 python bench.py --gpus 2 --steps 2 --warmup 1 --scaling strong > gpurun_out/bench_2gpu_strong.json 2> gpurun_out/bench_2gpu_strong.err; echo "strong rc=$?"; tail -2 gpurun_out/bench_2gpu_strong.err; cat gpurun_out/bench_2gpu_strong.json
timeout 600 python bench.py --gpus 2 --workload sweep --steps 1 --sweep-runs 8000 > gpurun_out/sweep_2gpu.json 2> gpurun_out/sweep_2gpu.err; echo "sweep rc=$?"; tail -2 gpurun_out/sweep_2gpu.err; cat gpurun_out/sweep_2gpu.json
