set -x
python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25
python tools/time_hpmf_growth.py 2 32 > gpurun_out/hpmf_growth_cfg2.json 2> gpurun_out/hpmf_growth.err; cat gpurun_out/hpmf_growth_cfg2.json; tail -3 gpurun_out/hpmf_growth.err
ncu --set full --clock-control none --import-source on -k regex:'sasa_kernel|hpmf_pair_kernel' -c 2 -o gpurun_out/hpmf_r1 -f python tools/time_hpmf.py 3 512 > gpurun_out/ncu_hpmf.log 2>&1; tail -3 gpurun_out/ncu_hpmf.log
