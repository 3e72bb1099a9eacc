timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python tools/profile_train_ops.py --batch 8 2>/dev/null | grep -E "attn_mqa_bwd|== backward|== forward" | head
timeout 300 python tools/kernel_sweep.py --no-torch 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    if d.get('core_kernel') == 'dv_attn_temporal': print(d['T'], d['H'], round(d['core_ms'], 3), round(d['core_GBps']))
" | head -4
