cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
mkdir -p /tmp/prof
prof() {  # name, kernel regex, title, command...
  local name=$1 rx=$2 title=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "$name plain run failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o /tmp/prof/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
  ncu -i /tmp/prof/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/prof/$name.ncu-rep gpurun_out/${name}_summary.md "$title" > /dev/null 2>&1
  rm -f /tmp/prof/$name.ncu-rep
}
prof r2_cnvm_batch cnvm_native "round 2 (final): cnvm_native_kernel<2,smem,csr,noalias,coop> on a batch of 2368 per-run ER graphs (N=1e5, 64-byte agent records, 4 warps per chain), t_max=2" python tools/probe_batch_kernel.py 2368 2 1
cat gpurun_out/plain_r2_cnvm_batch.log

