cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_20k_r1g.json 2> gpurun_out/bench_20k_r1g.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_20k_r1g.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['stages'], d['clocks'], d['cpu_baseline']['value'], d['roofline']['achieved'], d['roofline']['frac'])"
