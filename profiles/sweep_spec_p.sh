#!/bin/bash
# Sweep the lanes-per-agent variants of the specialised kernel on the bench workload.
for P in 8 4 2 1; do
  GSM_SPEC_P=$P python bench.py --steps 5000 --warmup 50 --no-cpu --e2e-steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('P=$P value %.3e  us/step %.2f  roofline %.3f' % (d['value'], d['roofline']['launch_us'], d['roofline']['frac']))"
done
