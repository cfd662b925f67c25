#!/bin/bash
python -m pytest tests/test_gpu_m_packed_ops.py tests/test_gpu_e_generator.py tests/test_gpu_n_testpair.py tests/test_gpu_c_conv.py tests/test_gpu_d_chain.py -m gpu -x -q 2>&1 | tail -6
python tools/profile_step.py 32 > gpurun_out/profile_step_s3h.txt 2>&1; head -12 gpurun_out/profile_step_s3h.txt | tail -10 | cut -c1-120
python bench.py --no-ops --train-steps 0 > gpurun_out/bench_s3h.json 2> gpurun_out/bench_s3h.err; tail -3 gpurun_out/bench_s3h.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_s3h.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['eager'], d['clocks'], d['gpu_launches'], d['bf16_mode']['value'], d['parity_vs_cpu_oracle'], d['roofline']['frac'], d['roofline']['traffic'], d['batch1'])"
