cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for a in "32 728 1" "32 1536 1" "64 728 1"; do BD_FUSE_SEPCONV=0 python tools/prof_dw.py $a 2>&1 | tail -1; done
python tools/op_table.py 16 8 > gpurun_out/op_table_r1r.txt 2>&1
grep "====" gpurun_out/op_table_r1r.txt
