#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_integrators.py -x -q -m gpu -k "light or triple or cornell" 2>&1 | tail -4
