set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_replay.py tests/test_gpu_native.py -m gpu -q 2>&1 | tail -40
