ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'shapelet_fwd_tc|shapelet_bwd_tc' --csv python tools/profile_step_kernels.py cosine 2>/dev/null | grep "shapelet_" | python -c '
import sys,csv
for r in csv.reader(sys.stdin):
    print(r[4][:60], r[-3], r[-1])'
python bench.py --no_extras --no_cpu_baseline --distance_func cosine --precision 3xtf32 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"])'
