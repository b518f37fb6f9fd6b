cd $GRAFT_REPO_ROOT
timeout 300 python tools/stress.py 1500 2>&1 | grep -v "Warning" | tail -1 | cut -c1-200
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/op_table.py 16 40 > gpurun_out/op_table_r1w.txt 2>&1; grep "====" gpurun_out/op_table_r1w.txt
