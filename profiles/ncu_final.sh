# final round-2 captures with the shipped code (pair plans launch the kernel once per step: no warm-up pass)
cd $GRAFT_REPO_ROOT
B="python bench.py --steps 3 --warmup 3 --min-warm-seconds 0 --no-cpu-baseline --no-lbph --no-c4 --no-c5"
$B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/ncu_bench.log 2>&1
for t in tc tc16; do
  python profiles/run_ncu_targets.py $t > gpurun_out/plain_$t.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cosine_tc -s 2 -c 1 -f -o gpurun_out/prof_${t}_r2 python profiles/run_ncu_targets.py $t > gpurun_out/ncu_$t.log 2>&1
  tail -1 gpurun_out/plain_$t.log
done
mkdir -p gpurun_out/summ
FRB_SUMMARY_OUT=gpurun_out/summ python profiles/summarize.py r2
find gpurun_out -name "*.ncu-rep" -size +8M -delete
ls -la gpurun_out/summ; du -sh gpurun_out
