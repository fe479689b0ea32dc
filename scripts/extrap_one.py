import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from extrap_probe import case
print("N=1025 1 disc R=164 1 layer: %.3f ms" % case(1025, [(0.5, 0.5)], 164, layers=1, reps=1))
