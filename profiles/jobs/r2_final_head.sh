#!/bin/bash
# Final validation of HEAD: GPU test suite, smoke, default bench.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/final/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/final/pytest.log
tail -3 gpurun_out/final/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/final/smoke.log
timeout 600 python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err; echo "bench rc=$?"
cat gpurun_out/final/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_ref.json 2>> gpurun_out/final/bench.err; echo "ref rc=$?"
cat gpurun_out/final/bench_ref.json
