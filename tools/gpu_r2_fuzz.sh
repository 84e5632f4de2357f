cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
fail=0
for o in $(seq ${FUZZ_FROM:-8} ${FUZZ_TO:-30}); do
  r=$(SPONET_FUZZ_OFFSET=$o timeout 600 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -1)
  echo "offset $o: $r"
done
