cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_graphgen.py -x -q 2>&1 | tail -2
timeout 1200 python bench.py --workload networks --steps 3 --warmup 3 > gpurun_out/r2_networks_1gpu.json 2> gpurun_out/r2_networks_1gpu.err; echo "rc=$?"
tail -3 gpurun_out/r2_networks_1gpu.err; cut -c1-1500 gpurun_out/r2_networks_1gpu.json
