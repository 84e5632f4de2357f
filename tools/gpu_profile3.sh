cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
export SPONET_B200_LIB=$PWD/sponet_b200/_variants/lib_p8a32.so
export SP_WPC=2
PCMD="python tools/perf_probe.py --model cnvm --t-max 1 --reps 2"
$PCMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cnvm_native -s 1 -c 1 -o gpurun_out/prof_cnvm_pair $PCMD > gpurun_out/ncu_full3.log 2>&1
echo "full rc=$?"
