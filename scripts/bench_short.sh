#!/bin/bash
# device-resident leg only, one line per option set:  scripts/bench_short.sh "" "--native-opt fused_stages=4" ...
for o in "$@"; do
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --wire-frames 0 --c4-frames 0 $o 2>gpurun_out/bench_short.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$o', '| frames/s %.0f ms/step %.3f frac %.3f kernel ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))" || tail -3 gpurun_out/bench_short.err
done
