set -x
timeout 600 python -m pytest tests/test_gpu_multi_area.py -x -q 2>&1 | tail -2
timeout 600 python scripts/configs_report.py 2>&1 | grep -A2 "multi_area_demo\|detailed_mc_2e6" | grep "kernel_ms\|years_per_s\|multi_area\|detailed"
