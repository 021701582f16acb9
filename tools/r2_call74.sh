timeout 300 python tools/small_phases.py 2>&1 | grep "cycles"
