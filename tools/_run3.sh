cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export TRACE_LINES=50 TRACE_SKIP=600
python tools/umma_trace.py 512 32 64 1 > gpurun_out/t9_32_64_1x1.txt 2>&1
python tools/umma_trace.py 512 64 64 3 > gpurun_out/t9_64_3x3.txt 2>&1
python tools/op_table.py 16 40 > gpurun_out/op_table_r1n.txt 2>&1
tail -12 gpurun_out/op_table_r1n.txt
