cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 900 python bench.py --gpus 8 --workload sweep --steps 2 --sweep-runs 100000 > gpurun_out/r2_sweep_8gpu.json 2> gpurun_out/r2_sweep_8gpu.err; echo "sweep rc=$?"
grep "^{" gpurun_out/r2_sweep_8gpu.json | cut -c1-300
