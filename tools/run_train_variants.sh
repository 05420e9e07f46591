#!/bin/bash
# usage: tools/run_train_variants.sh name1 name2 ...  -> training step time per variant library (Li-like and binary sets)
for v in "$@"; do
  for cfg in li bin; do
    NNMD_B200_LIB=$PWD/tools/variants/libnnmd_$v.so timeout 200 python tools/bench_train.py --config $cfg --frames 10000 --steps 40 --warmup 10 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', '$cfg', 'ms/step %.4f' % d['ms_per_step'], 'frames/s %.3fM' % (d['frames_per_s']/1e6))"
  done
done
