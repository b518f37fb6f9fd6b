cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --scene 5912 --steps 2 --warmup 2 > gpurun_out/bench_2gpu_r1d.json 2> gpurun_out/bench_2gpu_r1d.err
tail -c 1200 gpurun_out/bench_2gpu_r1d.json; tail -3 gpurun_out/bench_2gpu_r1d.err
