#!/bin/bash
# A/B of environment-selected launch variants with the precise timer (bench.py graph replay):
#   tools/ab_env.sh "VAR=1 OTHER=2" "" ...   ("" = defaults); workloads from $WLS (default: c2 c2hd c3)
WLS=${WLS:-"c2 c2hd c3"}
for rep in 1 2; do
  for v in "$@"; do
    for wl in $WLS; do
      env $v python bench.py --no-extra --steps 3000 --warmup 30 --workload $wl 2>/dev/null | \
        python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[%s]' % '$v', '$wl', 'us/step %.2f' % (d['ms_per_step']*1e3), 'frac %.3f' % d['roofline']['frac'])"
    done
  done
done
