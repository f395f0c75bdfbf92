timeout 200 python tools/sweep_quick.py 29 11 2>/dev/null > /tmp/a.txt; head -4 /tmp/a.txt
JDSP_FFT_BIG4K=1 timeout 200 python tools/sweep_quick.py 29 12 2>/dev/null > /tmp/b.txt; head -1 /tmp/b.txt
JDSP_FFT_NO_PIPE=1 timeout 200 python tools/sweep_quick.py 29 11 2>/dev/null > /tmp/c.txt; head -2 /tmp/c.txt
timeout 200 python tools/sweep_quick.py 27 11 2>/dev/null > /tmp/d.txt; head -4 /tmp/d.txt
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,power.limit,clocks_event_reasons.active --format=csv
