cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
for lib in "" sponet_b200/_variants/lib_nbpred.so; do
  export SPONET_B200_LIB=$lib; [ -z "$lib" ] && unset SPONET_B200_LIB
  timeout 600 python tools/probe_multi.py 100000:10000:10:101:0 10000:10000:10:100:0 100000:1250:10:101:0 2>&1 | tail -4
done
