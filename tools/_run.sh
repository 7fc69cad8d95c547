python tools/step_phases.py --tag polar_rows --steps 6 2>&1 | tail -1 > gpurun_out/r03_ab4.txt
cat gpurun_out/r03_ab4.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
