cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python tools/probe_multi.py 100000:10000:10:101:0 100000:1184:10:101:2 10000:10000:10:100:0 sis:1000000:2368:1:11:0 cntm:100000:10000:2:11:0
