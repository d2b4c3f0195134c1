timeout 600 python profiles/tools/smoke_margin_probe.py 2>&1 | grep -v Warning | tail -4
