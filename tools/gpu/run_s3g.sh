#!/bin/bash
# one GPU-box job of this session: graph-replay bench, smoke launch split, ncu of the dominant SPADE GEMM, memcheck of the new FIR kernels
python bench.py --no-ops --train-steps 0 > gpurun_out/bench_s3g.json 2> gpurun_out/bench_s3g.err; tail -3 gpurun_out/bench_s3g.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_s3g.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['eager'], d['clocks'], d['gpu_launches'], d['bf16_mode']['value'], d['parity_vs_cpu_oracle'])"
python tools/smoke_launches.py 2>&1 | tail -4
python -m pytest tests/test_gpu_e_generator.py tests/test_gpu_n_testpair.py -m gpu -x -q 2>&1 | tail -3
python tools/profile_step.py 32 > gpurun_out/profile_step_s3g.txt 2>&1
ncu --set full --clock-control none -k regex:igemm -s 2 -c 1 -o /tmp/prof_spade128 -f python tools/spade_layer.py 128 256 32 bf16x2 > gpurun_out/ncu_spade128_s3.log 2>&1
python tools/ncu_reduce.py /tmp/prof_spade128.ncu-rep gpurun_out/ncu_spade128_s3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_b_upfirdn2d.py tests/test_gpu_m_packed_ops.py -m gpu -x -q > gpurun_out/memcheck_s3.txt 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/memcheck_s3.txt
