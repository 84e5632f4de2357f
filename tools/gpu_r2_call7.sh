cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
for o in 1 2 3 4 5 6; do SPONET_FUZZ_OFFSET=$o timeout 600 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_native.py tests/test_gpu_sweep.py tests/test_gpu_statistics.py -m gpu -q -x 2>&1 | tail -2
