#!/usr/bin/env bash
# Whole-round check on one box: every GPU test, smoke(), the default bench line (both arms), then the ncu evidence.
tag="${1:-r2z}"; out=$PWD/gpurun_out; mkdir -p $out
python -m pytest tests -x -q -m gpu > $out/${tag}_all_tests.log 2>&1
echo "all tests rc=$?"; tail -3 $out/${tag}_all_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/${tag}_smoke.log
( time python bench.py --impl reference --steps 20 --warmup 5 ) > $out/${tag}_ref.json 2> $out/${tag}_bench.err
( time python bench.py --steps 20 --warmup 5 ) > $out/${tag}_bench.json 2>> $out/${tag}_bench.err
echo "bench rc=$?"; grep real $out/${tag}_bench.err
bash tools/gpu_profile_r2.sh ${tag}p
