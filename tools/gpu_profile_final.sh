# ncu evidence for the dominant kernel of bench.py: launch list of the bench command + one full capture
# of the SAME launch configuration (H, R=1e4, t_max=100, T=101).  Plain runs first.
cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
PCMD="python tools/perf_probe.py --model cnvm --t-max 100 --T 101 --reps 2"
$PCMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cnvm_native -s 1 -c 1 -o gpurun_out/prof_cnvm_final $PCMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -2 gpurun_out/plain2.log
