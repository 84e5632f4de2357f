# ncu capture of the one-warp-per-chain kernel (config-2 shape: N=1e4)
cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
PCMD="python tools/perf_probe.py --model cnvm --n 10000 --runs 20000 --t-max 10 --T 101 --reps 2"
$PCMD > gpurun_out/plain_n1e4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cnvm_native -s 7 -c 1 -o gpurun_out/prof_n1e4 $PCMD > gpurun_out/ncu_n1e4.log 2>&1
echo "full rc=$?"; tail -1 gpurun_out/plain_n1e4.log
