cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_states.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -6
for v in fa fb fc fd fe; do SPONET_B200_LIB=$PWD/sponet_b200/_variants/lib_$v.so python tools/probe_multi.py 100000:10000:10:101:0 100000:1184:10:101:2 2>&1 | grep -v "^$"; done
