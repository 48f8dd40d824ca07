set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/r01b_bench.json 2> gpurun_out/r01b_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r01b_ref.json 2>> gpurun_out/r01b_bench.err
python tools/kbench.py --slides 100 --patches 50000 --classes 30 > gpurun_out/r01b_k30.log 2>&1
python bench.py --slides 200 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --slides 200 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_keys_regw -s 3 -c 1 -o gpurun_out/r01b_score -f python bench.py --slides 200 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:head_rows_tc -s 3 -c 1 -o gpurun_out/r01b_head -f python bench.py --slides 200 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:select_mark -s 3 -c 1 -o gpurun_out/r01b_select -f python bench.py --slides 200 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_sel.log 2>&1
tail -2 gpurun_out/r01b_k30.log; cat gpurun_out/r01b_bench.json | cut -c1-600
