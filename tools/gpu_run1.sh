set -x
cd $GRAFT_REPO_ROOT
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02a_tests.log
timeout 600 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02a_ref.json 2>/dev/null
timeout 300 ncu --metrics $M --clock-control none -k regex:step_kernel -s 300 -c 8 --csv --log-file gpurun_out/r02a_counts_4096.csv python bench.py --steps 4 --warmup 3 --no-extras > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:step_kernel -s 300 -c 8 --csv --log-file gpurun_out/r02a_counts_65536.csv python bench.py --steps 4 --warmup 3 --no-extras --envs-per-gpu 65536 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 300 -c 1 -f -o gpurun_out/r02a_step_4096 python bench.py --steps 4 --warmup 3 --no-extras > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 300 -c 1 -f -o gpurun_out/r02a_step_65536 python bench.py --steps 4 --warmup 3 --no-extras --envs-per-gpu 65536 > /dev/null 2>&1
tail -3 gpurun_out/r02a_tests.log; cat gpurun_out/r02a_bench.json | head -c 1500
