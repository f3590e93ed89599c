cd $GRAFT_REPO_ROOT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:perform_tc_kernel -s 60 -c 1 -f -o gpurun_out/r02d_perform_tc python tools/dev_perform_tc.py > /dev/null 2>&1
ls -la gpurun_out | grep r02d
