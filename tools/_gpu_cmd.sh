set -x
ncu --set full --clock-control none --import-source on -k regex:'hpmf_trial_kernel' -s 150 -c 2 -o gpurun_out/hpmf_trial_r1 -f python tools/time_hpmf_growth.py 3 256 > gpurun_out/ncu_hpmf_trial.log 2>&1; tail -3 gpurun_out/ncu_hpmf_trial.log
