cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/op_table.py 16 8 > gpurun_out/op_table_r1q.txt 2>&1
grep "====" gpurun_out/op_table_r1q.txt
python bench.py --scene 5912 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_5912_r1q.json 2> gpurun_out/bench_5912_r1q.err; tail -c 600 gpurun_out/bench_5912_r1q.json
ncu --set full --clock-control none --import-source on -k regex:dwconv -s 2 -c 1 -o gpurun_out/prof_dw_r1q python tools/prof_dw.py 32 728 1 > gpurun_out/ncu_dw.log 2>&1
