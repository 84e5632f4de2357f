cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
H="100000:10000:10:101:0"; C2="10000:10000:10:100:0"
python tools/probe_multi.py $H $C2 100000:1184:10:101:2 100000:592:10:101:4 100000:296:10:101:8 10000:2368:10:100:1 100000:1250:10:101:0 \
   sis:1000000:2368:1:11:0 sis:1000000:148:1:11:0
export SPONET_B200_LIB=$PWD/sponet_b200/_variants/lib_t768.so
timeout 900 python -m pytest tests/test_gpu_native.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -4
python tools/probe_multi.py $H $C2 100000:1184:10:101:3 100000:1184:10:101:2 100000:1332:10:101:2 100000:592:10:101:6 10000:3552:10:100:1 10000:2368:10:100:1 100000:1250:10:101:0 \
   sis:1000000:2368:1:11:0 sis:1000000:148:1:11:0
