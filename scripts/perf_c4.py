import sys
sys.path.insert(0, '.')
from scripts.perf_probe import run
run(128, 16, 8, 4, P=512, eig=2)
run(128, 16, 8, 4, P=2048, eig=2, reps=1)
