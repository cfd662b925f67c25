#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; tail -3 gpurun_out/pytest_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_final.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['eager'], d['clocks'], d['gpu_launches'], d['bf16_mode']['value'], d['parity_vs_cpu_oracle'], d['roofline']['frac'], d['roofline']['launch'], d['roofline']['traffic'], d['batch1'], (d['train'] or {}).get('ms_per_step'))"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null; cat gpurun_out/bench_final_ref.json | cut -c1-400
python tools/profile_step.py 32 > gpurun_out/profile_step_final.txt 2>&1
python tools/gen_trace.py > gpurun_out/gen_trace_final.txt 2>&1
