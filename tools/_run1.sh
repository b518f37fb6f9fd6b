set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export TRACE_LINES=90 TRACE_SKIP=600
python tools/umma_trace.py 512 64 64 3 > gpurun_out/t2_64_3x3.txt 2>&1
python tools/umma_trace.py 256 64 256 1 > gpurun_out/t2_64_256_1x1.txt 2>&1
python tools/umma_trace.py 32 728 728 1 > gpurun_out/t2_728_1x1.txt 2>&1
python tools/umma_trace.py 256 32 32 3 > gpurun_out/t2_32_3x3.txt 2>&1
python tools/umma_trace.py 512 128 64 3 > gpurun_out/t2_128_64_3x3.txt 2>&1
python tools/prof_dw.py 32 728 1 > gpurun_out/dw_plain.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:dwconv -s 2 -c 1 -o gpurun_out/prof_dw_r1b python tools/prof_dw.py 32 728 1 > gpurun_out/ncu_dw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -o gpurun_out/prof_pw728_r1b python tools/prof_dw.py 32 728 1 > gpurun_out/ncu_pw.log 2>&1
tail -3 gpurun_out/t2_*.txt gpurun_out/dw_plain.txt
