cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_graphgen.py tests/test_gpu_native.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -2
timeout 300 python tools/probe_multi.py sis:1000000:2368:1:11:0 sis:1000000:148:1:11:0 100000:10000:10:101:0 h:100000:1184:10:101:0:g 2>&1 | tail -4
timeout 300 python tools/probe_batch_kernel.py 2368 4 1 0 2>&1 | tail -1
timeout 300 python tools/probe_batch_kernel.py 2368 4 0 0 2>&1 | tail -1
