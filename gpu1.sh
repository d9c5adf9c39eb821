for m in 512 576 640; do
echo "MAXT=$m depth 7"
PM4_MAXT=$m python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --max-tree-depth 7 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l); print(d['value'], d['ms_per_step'])
    except Exception: pass"
done
