cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
P="100000:10000:10:101:0 10000:10000:10:100:0 10000:2368:10:100:1 sis:1000000:2368:1:11:0 100000:1250:10:101:0"
python tools/probe_multi.py $P
SPONET_B200_LIB=$PWD/sponet_b200/_variants/lib_boolflags.so python tools/probe_multi.py $P
python tools/probe_multi.py $P
