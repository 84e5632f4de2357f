cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python tools/probe_multi.py 100000:10000:10:101:0 10000:10000:10:100:0 sis:1000000:2368:1:11:0 100000:100000:0.5:6:0
timeout 600 python bench.py --workload sweep --steps 2 --sweep-runs 100000 > gpurun_out/r2_sweep_1gpu.json 2> gpurun_out/r2_sweep_1gpu.err; echo "sweep rc=$?"; cat gpurun_out/r2_sweep_1gpu.json | cut -c1-400
