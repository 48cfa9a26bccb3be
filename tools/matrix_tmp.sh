python -m pytest tests/test_gpu_hpmf.py -q 2>&1 | tail -2
python tools/time_hpmf_fused.py cfg3 256,2048 2>&1 | grep "C="
python tools/time_hpmf_fused.py cfg2 128 2>&1 | grep "C="
python tools/time_hpmf_fused.py cfg4 512 2>&1 | grep "C="
