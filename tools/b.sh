#!/bin/bash
# usage: tools/b.sh [env...]  -> prints value ms_per_step launch_ms frac share
env "$@" timeout 300 python bench.py --scale 0.05 --steps 2 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3e bases/s  step %.2f ms  kernel %.2f ms  frac %.4f  share %.2f' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['roofline']['kernel_share_of_step']))"
