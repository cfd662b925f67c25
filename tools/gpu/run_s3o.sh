#!/bin/bash
python -m pytest tests/test_gpu_d_chain.py tests/test_gpu_c_conv.py tests/test_gpu_e_generator.py tests/test_gpu_n_testpair.py -m gpu -x -q 2>&1 | tail -6
python tools/ab_flag.py networks.UP2_PHASES 8 2>&1 | tail -2
