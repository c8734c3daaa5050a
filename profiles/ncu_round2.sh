set -x
cd $GRAFT_REPO_ROOT
B="python bench.py --steps 3 --warmup 3 --min-warm-seconds 0 --no-cpu-baseline --no-lbph --no-c4 --no-c5"
$B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/ncu_bench.log 2>&1
for t in "tc cosine_tc 4 2" "filter chisq_filter_kernel 2 1" "chisq_q1_u16 chisq_kernel 2 1" "chisq_q1_u8 chisq_kernel 2 1" "chisq_b_u8 chisq_kernel 2 1" "lbp lbp_hist_pipe_kernel 2 1" "lbp8 lbp_hist_pipe_kernel 2 1"; do
  set -- $t
  python profiles/run_ncu_targets.py $1 > gpurun_out/plain_$1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/prof_${1}_r2 python profiles/run_ncu_targets.py $1 > gpurun_out/ncu_$1.log 2>&1
  tail -2 gpurun_out/plain_$1.log
done
mkdir -p gpurun_out/summ
FRB_SUMMARY_OUT=gpurun_out/summ python profiles/summarize.py r2
ls -la gpurun_out/*.ncu-rep gpurun_out/summ
# only gpurun_out/ (<= 64 MiB) travels back: the text summaries always, the reports while they fit
find gpurun_out -name "*.ncu-rep" -size +8M -delete
du -sh gpurun_out
