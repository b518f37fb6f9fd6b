cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/op_table.py 16 8 > gpurun_out/op_table_r1s.txt 2>&1
grep "====" gpurun_out/op_table_r1s.txt
