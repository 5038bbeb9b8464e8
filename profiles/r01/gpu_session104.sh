#!/bin/bash
# mean-field auto-guides on the CUDA model
set -x
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "mean_field or elbo_terms" 2>&1 | tail -12
