for v in 1 0 1 0; do
if [ $v = 1 ]; then export PM4_NO_LOCKSTEP_LPT=1; else unset PM4_NO_LOCKSTEP_LPT; fi
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l); print('nolpt=$v', d['value'], d['ms_per_step'], d['clocks'])
    except Exception: pass"
done
