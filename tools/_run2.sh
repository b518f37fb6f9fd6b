cd $GRAFT_REPO_ROOT
export TRACE_LINES=120 TRACE_SKIP=900
python tools/umma_trace.py 512 64 64 3 > gpurun_out/t5_64_3x3.txt 2>&1
