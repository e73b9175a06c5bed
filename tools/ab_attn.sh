# A/B of differently built libevt.so files on the attention probe (development aid).
#   usage: bash tools/ab_attn.sh path/to/libevt_a.so path/to/libevt_b.so ...
for i in 1 2; do
  for v in "$@"; do
    echo -n "$v: "; EVT_LIB_PATH=$v python tools/attn_probe.py 1024 2>&1 | tail -1
  done
  echo -n "in-tree: "; python tools/attn_probe.py 1024 2>&1 | tail -1
done
