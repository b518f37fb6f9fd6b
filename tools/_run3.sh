cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_post_gpu.py -x -q 2>&1 | tail -2
BD_POST_TIMING=1 timeout 500 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/post_timing5.json 2> gpurun_out/post_timing5.err
grep -i "walk\|bd_contours\] trace" gpurun_out/post_timing5.err | tail -5
python -c "
import json; d=json.loads(open('gpurun_out/post_timing5.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['stages'])"
