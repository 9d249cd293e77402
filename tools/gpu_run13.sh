cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_queries.py tests/test_facade.py tests/test_gpu_patches.py -m gpu -x -q 2>&1 | tail -30
