cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/stress.py 150 2>&1 | grep -v "mbarrier timeout" | tail -2
timeout 300 python tools/op_table.py 16 40 > gpurun_out/op_table_r1u.txt 2>&1; grep "====" gpurun_out/op_table_r1u.txt
