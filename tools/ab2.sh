cd $GRAFT_REPO_ROOT; export PYTHONPATH=$PWD
ARGS=${AB_ARGS:---model cnvm --t-max 5 --T 51 --reps 3}
for round in 1 2 3; do
for v in "$@"; do
  export SPONET_B200_LIB=$PWD/sponet_b200/_variants/lib_$v.so
  echo "$v $(python tools/perf_probe.py $ARGS 2>&1 | grep '^rep' | awk '{print $6}' | tr '\n' ' ')"
done
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv,noheader
