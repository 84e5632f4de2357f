cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
nproc; nvidia-smi --query-gpu=name --format=csv,noheader
python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
H="100000:10000:10:101:0"; C2="10000:10000:10:100:0"
for v in r1 ph7; do SPONET_B200_LIB=$PWD/sponet_b200/_variants/lib_$v.so python tools/probe_multi.py $H $C2 2>&1 | grep -v "^$"; done
python tools/probe_multi.py $H $C2 100000:1332:10:101:1 100000:1184:10:101:2 100000:740:10:101:3 100000:592:10:101:4 100000:296:10:101:8 100000:148:10:101:8 \
   10000:2368:10:100:1 10000:1184:10:100:2 10000:592:10:100:4 10000:296:10:100:8 100000:1250:10:101:0 100000:1250:10:101:1
