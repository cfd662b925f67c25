#!/bin/bash
python -m pytest tests/test_gpu_d_chain.py tests/test_gpu_m_packed_ops.py tests/test_gpu_e_generator.py -m gpu -x -q 2>&1 | tail -6
python tools/ab_flag.py networks.UP2_PHASES 4 2>&1 | tail -2
