cd $GRAFT_REPO_ROOT
export TRACE_LINES=150 TRACE_SKIP=0
python tools/umma_trace.py 32 728 728 1 > gpurun_out/t3_728_1x1.txt 2>&1
python tools/umma_trace.py 512 64 64 3 > gpurun_out/t3_64_3x3.txt 2>&1
