cd $GRAFT_REPO_ROOT; export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_native.py -m gpu -q -x 2>&1 | tail -3
for round in 1 2; do for v in "$@"; do for w in 1 2; do
  export SPONET_B200_LIB=$PWD/sponet_b200/_variants/lib_$v.so
  echo "$v wpc=$w $(SP_WPC=$w timeout 300 python tools/perf_probe.py --model cnvm --t-max 5 --T 51 --reps 3 2>&1 | grep '^rep' | awk '{print $6}' | tr '\n' ' ')"
done; done; done
