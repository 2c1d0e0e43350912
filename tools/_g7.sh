cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_manager_gpu.py -x -q 2>&1 | grep -v "^ERROR\|^WARNING" | tail -40
