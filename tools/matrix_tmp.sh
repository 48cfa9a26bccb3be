python -m pytest tests/test_gpu_hpmf.py -q -x -k "per_trial_increments" 2>&1 | grep -E "^E  |Error" | head -6
python -m pytest tests/test_gpu_hpmf.py -q 2>&1 | tail -4
