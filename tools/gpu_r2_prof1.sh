cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
mkdir -p /tmp/prof
prof() {  # name, probe spec, steps for the summary
  ncu --set full --clock-control none --import-source on -k regex:cnvm_native -s 2 -c 1 -f -o /tmp/prof/$1 python tools/probe_multi.py $2 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i /tmp/prof/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof/$1.ncu-rep --page source --csv --print-source sass > gpurun_out/$1_src.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/prof/$1.ncu-rep gpurun_out/$1_summary.md "$3" > /dev/null 2>&1
}
prof r2_h 100000:10000:10:101:0 "round 2: cnvm_native_kernel<2,smem,nbtab,noalias,coop> on H: R=1e4, t_max=10, T=101"
prof r2_sis sis:1000000:2368:1:11:0 "round 2: cnvm_native_kernel<1,smem,nbtab,noalias,coop> on config 4: WS N=1e6, 2368 realizations, t_max=1"
ls -la gpurun_out/
