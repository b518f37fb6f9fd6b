cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for a in "32 728 1" "256 128 1" "256 128 2" "128 256 1" "64 728 2" "256 64 1" "32 1536 1"; do python tools/prof_dw.py $a 2>&1 | tail -1; done
python tools/op_table.py 16 12 > gpurun_out/op_table_r1l.txt 2>&1
tail -12 gpurun_out/op_table_r1l.txt
