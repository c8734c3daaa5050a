cd $GRAFT_REPO_ROOT
t=tc16
python profiles/run_ncu_targets.py $t > gpurun_out/plain_$t.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cosine_tc -s 2 -c 1 -f -o gpurun_out/prof_${t}_r2 python profiles/run_ncu_targets.py $t > gpurun_out/ncu_$t.log 2>&1
tail -1 gpurun_out/plain_$t.log
mkdir -p gpurun_out/summ
FRB_SUMMARY_OUT=gpurun_out/summ python profiles/summarize.py r2
find gpurun_out -name "*.ncu-rep" -size +8M -delete
