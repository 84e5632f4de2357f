cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
for lib in "" sponet_b200/_variants/lib_sets4.so sponet_b200/_variants/lib_sets5.so; do
  export SPONET_B200_LIB=$lib; [ -z "$lib" ] && unset SPONET_B200_LIB
  echo "== lib: ${lib:-default}"
  timeout 300 python tools/probe_batch_kernel.py 1184 4 1 2>&1 | tail -1
  SPONET_L2_FETCH=32 timeout 300 python tools/probe_batch_kernel.py 1184 4 1 2>&1 | tail -1 | sed 's/^/L2_FETCH=32: /'
  SPONET_L2_FETCH=64 timeout 300 python tools/probe_batch_kernel.py 1184 4 1 2>&1 | tail -1 | sed 's/^/L2_FETCH=64: /'
  timeout 300 python tools/probe_batch_kernel.py 1184 4 0 2>&1 | tail -1
  timeout 300 python tools/probe_multi.py sis:1000000:2368:1:11:0 2>&1 | tail -1
  timeout 600 python -m pytest tests/test_gpu_graphgen.py -x -q -k "cnvm_chains" 2>&1 | tail -1
done
