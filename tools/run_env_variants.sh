#!/bin/bash
# usage: tools/run_env_variants.sh LIB "NAME=ENV=VAL,ENV2=VAL2" ...   (tuning library tools/variants/libnnmd_<LIB>.so with environment knobs)
lib=$1; shift
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  ( export NNMD_B200_LIB=$PWD/tools/variants/libnnmd_$lib.so
    IFS=','; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done; unset IFS
    timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err )
  python - "$name" <<PY
import json,sys
v=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/var_{v}.json") if l.startswith("{")][-1])
    s=d["roofline"]["stage_ms_per_step"]
    print(f"{v:14s} step {d['ms_per_step']:.2f} ms  angular {s['angular']:.2f} gather {s['gather']:.2f} nlist {s['nlist']:.2f} radial {s['radial']:.2f} mlp {s['mlp']:.2f} | E {d['checksum']['energy']} sumF {d['checksum']['sum_abs_F']:.6f}")
except Exception as e:
    print(v, "ERR", e, open(f"gpurun_out/var_{v}.err").read()[-800:])
PY
done
