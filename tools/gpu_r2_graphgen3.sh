cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 900 python tools/probe_networks.py 10000 100 2>&1 | grep "device batches\|kernel\|Error\|error" | head
