cd $GRAFT_REPO_ROOT
timeout 300 python tools/_dbg_sql.py 2>&1 | tail -40
timeout 1500 python -m pytest tests/test_toon_gpu.py tests/test_toon_tp_gpu.py tests/test_sql_sanitizer.py -q -m gpu 2>&1 | grep -v "^ERROR\|^WARNING" | tail -4
