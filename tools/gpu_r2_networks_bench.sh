cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 1200 python bench.py --workload networks --steps 3 --warmup 3 > gpurun_out/r2_networks_1gpu.json 2> gpurun_out/r2_networks_1gpu.err; echo "rc=$?"
cut -c1-200 gpurun_out/r2_networks_1gpu.json
timeout 900 python tools/probe_networks.py 10000 100 2>&1 | grep "device batches"
