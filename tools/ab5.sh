cd $GRAFT_REPO_ROOT; export PYTHONPATH=$PWD
for round in 1 2 3; do for w in 1 2; do
  echo "wpc=$w $(SP_WPC=$w timeout 300 python tools/perf_probe.py --model cnvm --t-max 10 --T 101 --reps 5 2>&1 | grep '^rep' | awk '{print $6}' | sed 's/ev.s=//' | tr '\n' ' ')"
done; done
