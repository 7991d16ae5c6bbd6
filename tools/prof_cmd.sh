# Profiling commands of the round (run under gpurun; outputs land in gpurun_out/).
set -x
cd $GRAFT_REPO_ROOT
# launch list of about one bench step (7B, micro-batch 4 to keep the run short): skip the model build and part
# of the warm-up, stop the application after the window
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 --launch-count 6000 --kill on --csv \
    --log-file gpurun_out/r01_launches_bench7b_s2.csv python bench.py --steps 1 --warmup 3 --micro-batch 4 --skip-cpu-baseline \
    > gpurun_out/r01_launches_bench_s2.log 2>&1
ls -la gpurun_out/
