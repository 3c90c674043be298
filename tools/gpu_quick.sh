#!/bin/bash
# scratch visit to the GPU box used during development: the parity suite exactly as the driver runs it, smoke, and a short bench
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/driver_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/driver_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; echo "bench exit $?"
python tools/bench_brief.py gpurun_out/quick_bench.json
