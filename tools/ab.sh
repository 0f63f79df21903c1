#!/bin/bash
# A/B of library builds with the precise timer (bench.py graph replay, thousands of steps per event pair):
#   tools/ab.sh libA.so libB.so ...   (paths relative to the package dir; "-" = the default library)
P=$PWD/sports-field-homography_b200
for rep in 1 2; do
  for l in "$@"; do
    if [ "$l" = "-" ]; then unset SFH_LIB_PATH; else export SFH_LIB_PATH=$P/$l; fi
    for wl in c2 c2hd; do
      python bench.py --no-extra --steps 3000 --warmup 30 --workload $wl 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$l', '$wl', 'us/step %.2f' % (d['ms_per_step']*1e3))"
    done
  done
done
