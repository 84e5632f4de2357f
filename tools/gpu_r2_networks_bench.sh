cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_graphgen.py -x -q 2>&1 | tail -2
timeout 300 python tools/probe_batch_kernel.py 2368 4 1 0 2>&1 | tail -1
timeout 1200 python bench.py --workload networks --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_networks_1gpu_b.json 2> gpurun_out/r2_networks_1gpu_b.err; echo "rc=$?"
tail -3 gpurun_out/r2_networks_1gpu_b.err; cut -c1-200 gpurun_out/r2_networks_1gpu_b.json
