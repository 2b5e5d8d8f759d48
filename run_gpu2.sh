SRLX_SCHED=0 timeout 300 python tools/phase_clocks.py 2>&1 | tail -1 | cut -c1-50
SRLX_SCHED=4 timeout 300 python tools/phase_clocks.py 2>&1 | tail -1 | cut -c1-50
