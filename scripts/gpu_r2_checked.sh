#!/bin/bash
# full GPU tests, then the checked build (device-side index assertions) over the sanitizer probe and the parity / dense tests
set -u
OUT=gpurun_out/${1:-r2chk}
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log ); tail -4 $OUT/pytest_gpu.log
( DBA_LIB=checked timeout 600 python scripts/sanitize_probe.py > $OUT/checked_probe.log 2>&1; echo "checked probe exit $?" >> $OUT/checked_probe.log ); tail -3 $OUT/checked_probe.log
( DBA_LIB=checked timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dense.py tests/test_gpu_loss.py -m gpu -q > $OUT/checked_pytest.log 2>&1; echo "checked pytest exit $?" >> $OUT/checked_pytest.log ); tail -3 $OUT/checked_pytest.log
grep -c "DBA_CHECK failed" $OUT/checked_probe.log $OUT/checked_pytest.log
