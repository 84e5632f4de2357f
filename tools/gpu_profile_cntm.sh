cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
PCMD="python tools/perf_probe.py --model cntm --t-max 1 --T 11 --reps 2"
$PCMD > gpurun_out/plain_cntm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cntm_native -s 1 -c 1 -o gpurun_out/prof_cntm $PCMD > gpurun_out/ncu_cntm.log 2>&1
echo "full rc=$?"; tail -1 gpurun_out/plain_cntm.log
