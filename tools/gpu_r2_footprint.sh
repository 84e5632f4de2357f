cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
for n in 50000 25000; do for w in 1 2 4 8; do PROBE_N=$n timeout 300 python tools/probe_batch_kernel.py 2368 8 1 $w 2>&1 | tail -1; done; done
