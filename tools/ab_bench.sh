# A/B of differently built libevt.so files on the headline bench's stage times (development aid).
#   usage: bash tools/ab_bench.sh path/to/libevt_a.so ...     (the in-tree library runs last in every round)
show() { python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print(round(d['value']), {k: round(v['ms_per_launch'], 4) for k, v in d['stages'].items() if k[0] != '_'})"; }
for i in 1 2; do
  for v in "$@"; do echo -n "$v: "; EVT_LIB_PATH=$v python bench.py --steps 6 --no-cpu-baseline 2>/dev/null | show; done
  echo -n "in-tree: "; python bench.py --steps 6 --no-cpu-baseline 2>/dev/null | show
done
