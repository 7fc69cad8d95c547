python -m pytest tests/test_gpu_short_range.py tests/test_gpu_parity_fullsize.py -x -q -m gpu 2>&1 | tail -3
python tools/step_phases.py --tag overlap --steps 6 2>&1 | tail -1 > gpurun_out/r03_ab5.txt
PSIM_OVERLAP=0 python tools/step_phases.py --tag overlap_off --steps 6 2>&1 | tail -1 >> gpurun_out/r03_ab5.txt
cat gpurun_out/r03_ab5.txt
