#!/bin/bash
python -m pytest tests/test_gpu_a_bias_act.py tests/test_gpu_j_discriminator.py tests/test_gpu_c_conv.py tests/test_gpu_k_nccl_train.py -m gpu -x -q 2>&1 | tail -4
python tools/train_trace.py > gpurun_out/train_trace_s3m.txt 2>&1; head -22 gpurun_out/train_trace_s3m.txt | tail -20 | cut -c1-150
