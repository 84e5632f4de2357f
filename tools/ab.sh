cd $GRAFT_REPO_ROOT; export PYTHONPATH=$PWD
for v in "$@"; do
  export SPONET_B200_LIB=$PWD/sponet_b200/_variants/lib_$v.so
  echo "== $v"
  python tools/perf_probe.py --model cnvm --t-max 5 --T 51 --reps 2 2>&1 | tail -1 | cut -c1-60
  python tools/perf_probe.py --model cnvm --n 10000 --t-max 10 --reps 2 2>&1 | tail -1 | cut -c1-60
  python tools/perf_probe.py --model complete --n 1000 --runs 100000 --t-max 10 --reps 2 2>&1 | tail -1 | cut -c1-60
done
