cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
N=$(nvidia-smi -L | wc -l)
timeout 900 python bench.py --gpus $N --workload networks --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_networks_${N}gpu.json 2> gpurun_out/r2_networks_${N}gpu.err; echo "rc=$?"
tail -2 gpurun_out/r2_networks_${N}gpu.err | cut -c1-300; grep "^{" gpurun_out/r2_networks_${N}gpu.json | cut -c1-260
