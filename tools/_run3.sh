cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_post_gpu.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_20k_r1e_quick.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_20k_r1e_quick.json').read()); print(d['value'], d['e2e']['value'], d['stages'])"
