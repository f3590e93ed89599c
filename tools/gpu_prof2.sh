cd $GRAFT_REPO_ROOT
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 258 -c 1 -f -o gpurun_out/r02c_step_C5 python tools/prof_cfg.py C5 16384 6 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 258 -c 1 -f -o gpurun_out/r02c_step_C4 python tools/prof_cfg.py C4 16384 6 > /dev/null 2>&1
ls -la gpurun_out | grep r02c
