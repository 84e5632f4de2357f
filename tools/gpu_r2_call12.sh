cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
P="100000:10000:10:101:0 100000:1332:10:101:0 100000:1250:10:101:0 10000:10000:10:100:0 sis:1000000:2368:1:11:0"
python tools/probe_multi.py $P
SPONET_B200_LIB=$PWD/sponet_b200/_variants/lib_t576.so python tools/probe_multi.py $P
SPONET_B200_LIB=$PWD/sponet_b200/_variants/lib_t576.so timeout 600 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_native.py -m gpu -q -x 2>&1 | tail -2
