#!/bin/bash
# Round-2 evidence run (one GPU): each ncu pass only after the same command exited 0 without ncu.
set -u
O=gpurun_out/r2ev; mkdir -p $O
python -m pytest tests/test_gpu_pipeline.py -x -q > $O/pytest_pipeline.log 2>&1; echo "pipeline tests rc $?"
python bench.py --steps 10 --warmup 3 --sweep-out $O/large_sweep.jsonl > $O/bench.json 2> $O/bench.err; echo "bench rc $?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum
python tools/prof_step.py --skip-cir > $O/cp_plain.log 2>&1 && \
ncu --metrics $M --clock-control none -c 400 --csv --log-file $O/launches_cp.csv python tools/prof_step.py --skip-cir > $O/cp_ncu.log 2>&1
echo "cp launches rc $?"
for R in 1250000 10000000 1000000; do
  python tools/prof_step.py --skip-cp --rows $R $( [ $R = 1000000 ] && echo --queries 4096 ) > $O/search_plain_$R.log 2>&1 && \
  ncu --metrics $M --clock-control none -c 60 --csv --log-file $O/launches_search_$R.csv python tools/prof_step.py --skip-cp --rows $R $( [ $R = 1000000 ] && echo --queries 4096 ) > $O/search_ncu_l_$R.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k 'regex:tc_kernel|merge_rerank|topk_merge' --launch-skip 6 -c 3 -o $O/search_$R -f python tools/prof_step.py --skip-cp --rows $R $( [ $R = 1000000 ] && echo --queries 4096 ) > $O/search_ncu_f_$R.log 2>&1
  echo "search $R rc $?"
done
python tools/time_ffn.py 82158 > $O/ffn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ffn_block --launch-skip 3 -c 1 -o $O/ffn_block -f python tools/time_ffn.py 82158 > $O/ffn_ncu.log 2>&1
echo "ffn rc $?"
python tools/time_ffn.py 82158 --ln > $O/ffn_ln_plain.log 2>&1
ls -la $O
