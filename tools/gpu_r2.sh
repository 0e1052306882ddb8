#!/usr/bin/env bash
# Round-2 GPU check: new tests first, then the default bench line, then the whole GPU suite.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2.sh TAG'
tag="${1:-r2}"
mkdir -p gpurun_out
python -m pytest tests/test_abi_hardening_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu > gpurun_out/${tag}_new_tests.log 2>&1
echo "new tests rc=$?" | tee -a gpurun_out/${tag}_new_tests.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${tag}_bench.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/${tag}_ref.json 2>> gpurun_out/${tag}_bench.err
if [ "${2:-}" = "all" ]; then
  python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_all_tests.log 2>&1
  echo "all tests rc=$?"; tail -3 gpurun_out/${tag}_all_tests.log
fi
tail -5 gpurun_out/${tag}_new_tests.log
