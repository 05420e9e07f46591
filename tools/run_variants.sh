#!/bin/bash
# usage: tools/run_variants.sh "bench args" name1 name2 ...  -> one line per variant (ms/step, angular ms)
args=$1; shift
for v in "$@"; do
  NNMD_B200_LIB=$PWD/tools/variants/libnnmd_$v.so python bench.py $args --no-cpu-baseline > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - "$v" <<'P'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/var_{v}.json').read().strip().splitlines()[-1])
    s=d['roofline']['stage_ms_per_step']
    print(v, 'ms/step %.2f'%d['ms_per_step'], 'angular %.2f gather %.2f nlist %.2f'%(s['angular'],s['gather'],s['nlist']), flush=True)
except Exception as e:
    print(v, 'FAILED', e, open(f'gpurun_out/var_{v}.err').read()[-400:])
P
done
