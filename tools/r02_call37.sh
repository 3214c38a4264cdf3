#!/bin/bash
# call 37: final state after the head changes -- full GPU suite, smoke, pipeline bench, config-4 end to end
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_c37_pytest.txt 2>&1
tail -4 gpurun_out/r02_c37_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/bench_pipeline.py > gpurun_out/r02_c37_pipeline.json 2> gpurun_out/r02_c37_pipeline.err
cut -c1-700 gpurun_out/r02_c37_pipeline.json; tail -2 gpurun_out/r02_c37_pipeline.err
timeout 600 python tools/e2e_config4.py --out gpurun_out/r02_c37_e2e_config4.json > gpurun_out/r02_c37_e2e.log 2>&1
grep -E "ms_per_frame|rows_identical|t_rel_err_vs_oracle_max|objects_per_s" gpurun_out/r02_c37_e2e.log
timeout 300 python tools/bench_head.py --rois 128 > gpurun_out/r02_c37_head128.json 2>/dev/null
cut -c1-250 gpurun_out/r02_c37_head128.json
