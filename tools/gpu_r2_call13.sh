cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python tools/probe_multi.py 100000:10000:10:101:0 100000:1332:10:101:0 100000:1250:10:101:0 100000:2500:10:101:0 100000:5000:10:101:0 10000:10000:10:100:0 sis:1000000:2368:1:11:0
SPONET_FUZZ_OFFSET=7 timeout 600 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -1
