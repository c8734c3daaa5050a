cd $GRAFT_REPO_ROOT
python profiles/run_ncu_targets.py tc > gpurun_out/plain_tc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cosine_tc -s 4 -c 2 -f -o gpurun_out/prof_tc_r2 python profiles/run_ncu_targets.py tc > gpurun_out/ncu_tc.log 2>&1
tail -2 gpurun_out/ncu_tc.log
mkdir -p gpurun_out/summ
FRB_SUMMARY_OUT=gpurun_out/summ python profiles/summarize.py r2
ls -la gpurun_out/summ gpurun_out/*.ncu-rep
