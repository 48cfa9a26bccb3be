ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hpmf|sasa' --csv --log-file gpurun_out/hpmf_launches.csv python tools/time_hpmf_growth.py 3 256 > gpurun_out/ncu_hl.log 2>&1
tail -2 gpurun_out/ncu_hl.log
