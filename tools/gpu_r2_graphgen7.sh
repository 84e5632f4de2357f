cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
for w in 1 2 4 8 16; do timeout 300 python tools/probe_batch_kernel.py 2368 4 1 $w 2>&1 | tail -1; done
for g in 148 296 592; do timeout 300 python tools/probe_batch_kernel.py $g 8 1 0 2>&1 | tail -1; done
