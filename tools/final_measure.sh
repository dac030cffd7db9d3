set -x
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for c in C3 C3b C1 C2 C4 C5; do
  cc=${c%b}
  python bench.py --config $cc > gpurun_out/r02f_$c.json 2> gpurun_out/r02f_$c.err
  tail -1 gpurun_out/r02f_$c.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$c', round(d['value'],1), round(d['e2e']['value'],1), r and round(r['frac'],3), r and round(r['in_timed_region']['frac'],3), round(d['pipeline_hbm_frac'],3), d['clocks']['reasons'], d['cpu_baseline']['value'], d['gpu_launches'])"
done
python bench.py --config C4 --lk-step 4 > gpurun_out/r02f_C4s4.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_C3_reference.json 2>/dev/null; tail -c 400 gpurun_out/r02f_C3_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 3 --frames-per-step 8 > /dev/null 2>&1
python tools/launch_share.py gpurun_out/r02f_launches.csv | tee gpurun_out/r02f_launches_bench_4k.txt
