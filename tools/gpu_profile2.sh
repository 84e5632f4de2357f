cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
PCMD="python tools/perf_probe.py --model cnvm --t-max 1 --reps 2"
$PCMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cnvm_native -s 1 -c 1 -o gpurun_out/prof_cnvm_v3 $PCMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
