cd $GRAFT_REPO_ROOT
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 2 --warmup 2 > gpurun_out/bench_4gpu_20k_r1e.json 2> gpurun_out/bench_4gpu_20k_r1e.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_4gpu_20k_r1e.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stages'])"
tail -2 gpurun_out/bench_4gpu_20k_r1e.err | cut -c1-300
