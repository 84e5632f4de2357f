cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_graphgen.py tests/test_gpu_api.py -x -q 2>&1 | tail -1
timeout 900 python tools/probe_networks.py 10000 100 2>&1 | grep "device batches"
timeout 900 python tools/probe_networks.py 2368 10 2>&1 | grep "device batches\|generate\|host"
