timeout 300 python -m pytest tests/test_gpu_b_conv.py -m gpu -q --no-header -rf -x -k "levels or groupnorm" 2>&1 | grep -v "^  " | tail -30 > gpurun_out/pytest_r02h.log; tail -3 gpurun_out/pytest_r02h.log
HN_CONV_TABLE=gpurun_out/conv_table_r02h.json timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_r02h.log 2> gpurun_out/bench_r02h.err; tail -c 300 gpurun_out/bench_r02h.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02h.log').read().strip().splitlines()[-1])
print(d['value'], d['sequential']['value'], d['e2e']['value'], d['roofline']['frac'])"
timeout 300 python tools/phase_timing.py 2>&1 | tail -14
