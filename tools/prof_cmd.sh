# Profiling commands of the round (run under gpurun; outputs land in gpurun_out/).
set -x
cd $GRAFT_REPO_ROOT
# (1) launch list of about one bench step (7B, micro-batch 4 to keep the serialised run short): skip the model
#     build and part of the warm-up, stop the application after the window
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 --launch-count 6000 --kill on --csv \
    --log-file gpurun_out/r01_launches_bench7b_s3.csv python bench.py --steps 1 --warmup 3 --micro-batch 4 --skip-cpu-baseline \
    > gpurun_out/r01_launches_bench_s3.log 2>&1
# (2) one full-set capture of the forward + backward chain (AVT, r = 16, d = 4096, T = 8192): dram bytes per kernel
ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 42 -c 14 -o gpurun_out/r03_prof_chain \
    python tools/microbench.py --ranks 16 --variants avt --iters 4 --warmup 2 > gpurun_out/r03_prof_chain.log 2>&1
ls -la gpurun_out/
