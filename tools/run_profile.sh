# Round profile run (on the GPU box, from the repo root): bench lines of every BASELINE config, the launch lists and
# one ncu --set full capture per hot kernel.  Outputs land in gpurun_out/<tag>_*; summarise with tools/ncu_summary.py.
TAG=${1:-r02}
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_ref_cfg2.json 2>> gpurun_out/${TAG}_bench.err
for wl in cfg3 cfg3bank cfg4 cfg5; do
  python bench.py --workload $wl --steps 10 --warmup 3 --cpu-seconds 8 > gpurun_out/${TAG}_bench_${wl}.json 2>> gpurun_out/${TAG}_bench.err
done
KREGEX='regex:score_keys|score_bank|select_mark|compact_kernel|head_rows|pool_final|cross_entropy|head_tc_prep|head_f16_prep'
python tests/kbench_heads.py > gpurun_out/${TAG}_heads.log 2>&1
python tools/time_loops.py > gpurun_out/${TAG}_time_loops.log 2>&1
SMALL="--slides 200 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
SMALL4="--workload cfg4 --slides 80 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
SMALLB="--workload cfg3bank --slides 200 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
FULL="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e"     # launch lists at the full size of the bench lines
python bench.py $FULL > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $FULL > gpurun_out/ncu_l.log 2>&1
python bench.py --workload cfg4 $FULL > gpurun_out/plain4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 60 --csv --log-file gpurun_out/${TAG}_launches_cfg4.csv python bench.py --workload cfg4 $FULL > gpurun_out/ncu_l4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_keys_regw -s 3 -c 1 -o gpurun_out/${TAG}_score -f python bench.py $SMALL > gpurun_out/ncu_s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:head_rows_f16t -s 3 -c 1 -o gpurun_out/${TAG}_head -f python bench.py $SMALL > gpurun_out/ncu_h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:select_mark -s 3 -c 1 -o gpurun_out/${TAG}_select -f python bench.py $SMALL > gpurun_out/ncu_sel.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_keys_tc -s 3 -c 1 -o gpurun_out/${TAG}_score_tc -f python bench.py $SMALL4 > gpurun_out/ncu_stc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:head_rows_f16t -s 3 -c 1 -o gpurun_out/${TAG}_head_c30 -f python bench.py $SMALL4 > gpurun_out/ncu_htc.log 2>&1
python bench.py $SMALLB > gpurun_out/plainb.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_bank_tc -s 3 -c 1 -o gpurun_out/${TAG}_score_bank -f python bench.py $SMALLB > gpurun_out/ncu_sb.log 2>&1
for f in gpurun_out/${TAG}_bench_*.json; do python -c "
import json,sys
d=json.load(open('$f')); r=d['roofline']; e=d.get('e2e') or {}; c=d.get('cpu_baseline') or {}
print('$f', round(d['value']), 'slides/s', round(d['ms_per_step'],2), 'ms', r['kernel'], round(r['achieved']), r['unit'], 'frac', round(r['frac'],3), 'e2e', round(e.get('value',0)), 'cpu', round(c.get('value',0)))"; done
