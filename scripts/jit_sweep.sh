for cfg in "2 256 8" "1 256 8" "4 256 8" "2 128 16" "2 512 4" "4 128 16" "2 256 4" "2 256 16" "1 512 4" "4 512 2"; do
  set -- $cfg
  r=$(B200_JIT_U=$1 B200_JIT_BLOCK=$2 B200_JIT_CTAS_PER_SM=$3 timeout 100 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-train 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['value'])")
  echo "U=$1 block=$2 ctas/sm=$3 -> $r"
done
