# ncu evidence of the step kernel (tag = $1): launch counts + one --set full capture at 4096 and 65536 envs
set -x
cd $GRAFT_REPO_ROOT
T=$1
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for n in 4096 65536; do
timeout 300 ncu --metrics $M --clock-control none -k regex:step_kernel -s 300 -c 8 --csv --log-file gpurun_out/${T}_counts_$n.csv python bench.py --steps 4 --warmup 3 --no-extras --envs-per-gpu $n > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 300 -c 1 -f -o gpurun_out/${T}_step_$n python bench.py --steps 4 --warmup 3 --no-extras --envs-per-gpu $n > /dev/null 2>&1
done
ls -la gpurun_out/ | tail -8
