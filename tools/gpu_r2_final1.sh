cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
python __graft_entry__.py --smoke 2>&1 | tail -2
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:cnvm_native -s 6 -c 2 --csv --log-file gpurun_out/r2_bench_dram.csv $BENCH > gpurun_out/r2_ncu_dram.log 2>&1
echo "dram rc=$?"
timeout 1200 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --workload sweep --steps 2 --sweep-runs 100000 > gpurun_out/r2_sweep_1gpu.json 2> gpurun_out/r2_sweep_1gpu.err; echo "sweep rc=$?"
