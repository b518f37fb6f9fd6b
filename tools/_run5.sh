cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_20k_r1e.json 2> gpurun_out/bench_20k_r1e.err
tail -c 400 gpurun_out/bench_20k_r1e.json
python tools/traffic_batch.py > /dev/null 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_umma --csv --log-file gpurun_out/traffic_r1e.csv python tools/traffic_batch.py > gpurun_out/ncu_traffic.log 2>&1
python tools/prof_conv.py 512 64 64 3 > gpurun_out/prof_conv_plain_e.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -o gpurun_out/prof_conv64_r1e python tools/prof_conv.py 512 64 64 3 > gpurun_out/ncu_c64.log 2>&1
cat gpurun_out/prof_conv_plain_e.txt
python bench.py --scene 1592 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1592_plain_e.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --scene 1592 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_e.log 2>&1
