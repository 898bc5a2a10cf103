"""One beam of N particles tracked over a range on the CUDA product (for ncu captures at sizes other than bench.py's):
   ncu ... python tools/prof_run.py 16777216 TBDB02 01QM10 [sc=1]"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import harness as H
from hpsim_b200 import HPSim
HPSim.print_timing = False
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
first = sys.argv[2] if len(sys.argv) > 2 else "TBDB02"
last = sys.argv[3] if len(sys.argv) > 3 else "01QM10"
sc = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
c = H.Case(H.cuda_lib(), n=n, sc=sc)
r = c.run(first, last)
print("PROF_RUN", n, first, last, r)
