#!/bin/bash
# round 2, GPU call 25: evidence for the reworked k_ising2d: ncu --set full of one half-sweep of the default shape (16 band
# launches) and of one whole-lattice launch (MCX_BANDS=0), the launch list of a short bench, and the default bench line
mkdir -p gpurun_out/r02
B="python bench.py --steps 1 --warmup 1 --sweeps-per-step 5 --no-pt --no-cpu --no-extras"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ising2d -s 96 -c 16 -f -o gpurun_out/r02/ncu_ising2d_v8_bands $B > gpurun_out/r02/call25_ncu_bands.log 2>&1
MCX_BANDS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ising2d -s 8 -c 2 -f -o gpurun_out/r02/ncu_ising2d_v8_plain $B > gpurun_out/r02/call25_ncu_plain.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02/launches_v8.csv python bench.py --steps 2 --warmup 1 --sweeps-per-step 10 --no-cpu --no-pt --no-extras > gpurun_out/r02/call25_ncu_launches.log 2>&1
( time timeout 1200 python bench.py ) > gpurun_out/r02/call25_bench.json 2> gpurun_out/r02/call25_bench.err
tail -3 gpurun_out/r02/call25_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02/call25_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'kernel', d['roofline']['kernel_attempts_per_ns'], 'e2e', d['e2e']['value'], 'pt', d.get('pt',{}).get('value'), d.get('pt_every_sweep',{}).get('value'))
print({k:(v.get('value') if isinstance(v,dict) else v) for k,v in d.get('configs',{}).items()})
"
ls -la gpurun_out/r02/ | grep -i "ncu_ising2d_v8\|launches_v8"
