#!/bin/bash
# quick A/B of the headline kernel: parity tests of the fused grid-head path, then the device-timed step (no extras)
python -m pytest tests/test_gpu_yolo.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-extras --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('us/step', round(d['ms_per_step']*1e3,2), 'Mimg/s', round(d['value']/1e6,2), 'e2e us', round(d['e2e']['ms_per_step']*1e3,2), 'frac', round(d['roofline']['frac'],4))"
