set -u
O=gpurun_out/r2; mkdir -p $O
timeout 200 python scripts/r2_debug.py > $O/debug.txt 2>&1; grep -c "match=True" $O/debug.txt; grep "match=False\|watchdog" $O/debug.txt | head
timeout 900 python -m pytest tests/test_gpu_resident2.py tests/test_gpu_resident.py -x -q > $O/pytest_r2.log 2>&1; echo "pytest rc=$? $(tail -3 $O/pytest_r2.log | tr '\n' ' ')"
