# final-code ncu evidence (tag = $1): counts + --set full at 4096 / 65536 envs and C4 / C5; summarised ON the box (the reports are
# too big for gpurun_out's 64 MiB), only the text summaries travel back
set -x
cd $GRAFT_REPO_ROOT
T=$1
bash tools/gpu_prof.sh $T > /dev/null 2>&1
for c in C4 C5; do
timeout 200 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 258 -c 1 -f -o gpurun_out/${T}_step_$c python tools/prof_cfg.py $c 16384 2 > /dev/null 2>&1
done
cp profiles/kernel_counts.json /tmp/kc_before.json
timeout 120 python tools/ncu_summary.py $T
cp profiles/${T}_* gpurun_out/ ; cp profiles/kernel_counts.json gpurun_out/${T}_kernel_counts.json
for c in 4096 65536 C4 C5; do
timeout 120 python tools/ncu_phase_breakdown.py gpurun_out/${T}_step_$c.ncu-rep > gpurun_out/${T}_phase_breakdown_$c.txt 2> /dev/null
done
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | grep $T
