cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_graphgen.py -x -q 2>&1 | tail -1
timeout 300 python tools/probe_batch_kernel.py 1184 4 1 2>&1 | tail -1
timeout 300 python tools/probe_batch_kernel.py 2368 4 1 2>&1 | tail -1
timeout 300 python tools/probe_batch_kernel.py 592 4 1 2>&1 | tail -1
timeout 300 python tools/probe_batch_kernel.py 1184 4 0 2>&1 | tail -1
timeout 300 python tools/probe_multi.py sis:1000000:2368:1:11:0 100000:10000:10:101:0 2>&1 | tail -2
