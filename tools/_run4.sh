cd $GRAFT_REPO_ROOT
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_20k_r1c.json 2> gpurun_out/bench_20k_r1c.err
tail -c 1500 gpurun_out/bench_20k_r1c.json
python bench.py --scene 1592 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1592_plain.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --scene 1592 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/prof_conv.py 512 64 64 3 > gpurun_out/prof_conv_plain.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -o gpurun_out/prof_conv64_r1c python tools/prof_conv.py 512 64 64 3 > gpurun_out/ncu_c64.log 2>&1
python tools/prof_conv.py 128 256 256 3 >> gpurun_out/prof_conv_plain.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -o gpurun_out/prof_conv256_r1c python tools/prof_conv.py 128 256 256 3 > gpurun_out/ncu_c256.log 2>&1
cat gpurun_out/prof_conv_plain.txt
