cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
M="--metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum --clock-control none -k regex:cnvm_native -s 2 -c 1 --csv"
for lib in "" sponet_b200/_variants/lib_ld64.so; do
  export SPONET_B200_LIB=$lib; [ -z "$lib" ] && unset SPONET_B200_LIB
  echo "== lib: ${lib:-default}"
  timeout 300 python tools/probe_batch_kernel.py 1184 4 1 2>&1 | tail -1
  ncu $M python tools/probe_batch_kernel.py 1184 4 1 2>&1 | grep -i "cnvm_native" | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"'
  SPONET_L2_FETCH=32 ncu $M python tools/probe_batch_kernel.py 1184 4 1 2>&1 | grep -i "cnvm_native" | awk -F'","' '{print "L2F32:", $(NF-2), $(NF)}' | tr -d '"'
done
