cd $GRAFT_REPO_ROOT
export BD_CONTOUR_ONE_WALK=1
timeout 600 python -m pytest tests/test_post_gpu.py -x -q 2>&1 | tail -3
BD_POST_TIMING=1 timeout 500 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/post_timing2.json 2> gpurun_out/post_timing2.err
grep -i "bd_contours" gpurun_out/post_timing2.err | tail -4
python -c "
import json; d=json.loads(open('gpurun_out/post_timing2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['stages'])"
