# Round-2 ncu evidence (one gpurun call, one GPU).  Every profiled command first runs plain and must exit 0.
# Reports are summarised ON the box (raw / source CSV + markdown) -- the .ncu-rep files are too large to travel.
cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
mkdir -p /tmp/prof
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs"
$BENCH > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err && echo "plain bench ok" &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv $BENCH > gpurun_out/r2_ncu_list.log 2>&1
echo "list rc=$?"
# DRAM traffic of the timed launch itself (t_max = 100): metrics-only pass, the 6th matching launch = first timed step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:cnvm_native -s 5 -c 1 --csv --log-file gpurun_out/r2_bench_dram.csv $BENCH > gpurun_out/r2_ncu_dram.log 2>&1
echo "dram rc=$?"
prof() {  # name, kernel regex, command..., title
  local name=$1 rx=$2 title=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "$name plain run failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o /tmp/prof/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
  ncu -i /tmp/prof/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/$name.ncu-rep --page source --csv --print-source sass > gpurun_out/${name}_sass.csv 2>/dev/null
  ncu -i /tmp/prof/$name.ncu-rep --page source --csv --print-source cuda > gpurun_out/${name}_cuda.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/prof/$name.ncu-rep gpurun_out/${name}_summary.md "$title" > /dev/null 2>&1
  rm -f /tmp/prof/$name.ncu-rep
}
prof r2_cnvm_h cnvm_native "round 2: cnvm_native_kernel<2,smem,nbtab,noalias,coop> on H: R=1e4, t_max=10, T=101" python tools/probe_multi.py 100000:10000:10:101:0
prof r2_cnvm_n1e6 cnvm_native "round 2: cnvm_native_kernel<1,smem,nbtab,noalias,coop16> on config 4: WS N=1e6, 2368 realizations, t_max=1" python tools/probe_multi.py sis:1000000:2368:1:11:0
prof r2_cntm cntm_native "round 2: cntm_native_kernel<smem> on config 3 shape: BA N=1e5 m=5, R=1e4, t_max=2" python tools/probe_multi.py cntm:100000:10000:2:11:0
prof r2_cnvm_n1e4 cnvm_native "round 2: cnvm_native_kernel<2,smem,nbtab,noalias> on config 2: ER N=1e4, R=1e4, t_max=10, T=100" python tools/probe_multi.py 10000:10000:10:100:0
ls -la gpurun_out | tail -30
