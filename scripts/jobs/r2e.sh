#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_gpus.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r2e_pytest.log
