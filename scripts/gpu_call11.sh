#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_best_multiexp_2_22_2_24 --deselect tests/test_gpu_parity.py::test_best_fft_2_22_2_24 --deselect tests/test_gpu_parity.py::test_domain_transforms_k20_k22 ) > gpurun_out/r2c11_pytest.log 2>&1
tail -n 6 gpurun_out/r2c11_pytest.log
