cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_graphgen.py tests/test_gpu_native.py -x -q 2>&1 | tail -4
timeout 600 python tools/probe_networks.py 2368 10 2>&1 | tail -9
