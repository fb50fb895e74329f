#!/bin/bash
# band host tables extracted into a device-free function: GPU parity of the banded / segmented layouts with the
# invariant check switched on inside band_build (GDN_BAND_CHECK=1)
mkdir -p gpurun_out
export PYTHONPATH=$PWD
GDN_BAND_CHECK=1 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "banded or segmented or resident" > gpurun_out/c39_pytest.log 2>&1; tail -4 gpurun_out/c39_pytest.log
