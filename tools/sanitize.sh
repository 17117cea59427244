#!/bin/bash
# compute-sanitizer memcheck over the small GPU parity tests (one tool per gpurun call, B200_PROFILING.md): every kernel of the
# library on small trees -- upload / device builder / live trees, descent kernels, fits, binned + direct + dense evaluation, extras.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 \
  python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -m gpu -q -x \
  -k "not full_size and not c5_highdim and not k3b and not c4 and not large_n and not committed" > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck exit code $?" >> gpurun_out/r2_memcheck.log
tail -15 gpurun_out/r2_memcheck.log
