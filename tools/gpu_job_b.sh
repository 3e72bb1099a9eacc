for b in 64 96 128; do timeout 600 python bench.py --batch $b --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print($b, round(d['value'], 1), round(d['ms_per_step'], 2), d['clocks']['sm_mhz'], round(d['e2e']['value'], 1))"; done
