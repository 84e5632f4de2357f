cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_graphgen.py -x -q 2>&1 | tail -15
timeout 600 python tools/probe_networks.py 2368 10 2>&1 | tail -16
