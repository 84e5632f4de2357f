# ncu evidence for the dominant kernel.  Plain run first (must exit 0), then the launch list and one full capture.
cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
PCMD="python tools/perf_probe.py --model cnvm --t-max 1 --reps 2"
$PCMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cnvm_native -s 1 -c 1 -o gpurun_out/prof_cnvm $PCMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
ls -la gpurun_out
