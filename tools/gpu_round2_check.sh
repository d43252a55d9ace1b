#!/bin/bash
# the whole GPU suite, the synchronous-caller table and the default bench line, in one call
TAG=${1:-r02g}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
bash tools/gpu_callers_ab.sh ${TAG}_callers > /dev/null 2>&1; grep -v phases gpurun_out/${TAG}_callers.txt | cut -c1-250
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; head -c 900 gpurun_out/${TAG}_bench.json
