"""Registers / spills of the hot kernels from the ptxas logs of the last build: python scripts/regs.py [name-filter ...]"""
import re, glob, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
flt = sys.argv[1:] or ["k_mesh_weights_interp", "k_interp", "k_near_search", "k_near_slow", "k_chain", "k_tail", "k_path_fix", "k_mesh_weights"]
for f in sorted(glob.glob(os.path.join(ROOT, "gonzag_b200/build/*.o.log"))):
    txt = open(f).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt, re.S):
        d = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(GzbGrid")[0]
        if any(k in d for k in flt):
            print("%-64s regs %3s  stack %3s  spill %s/%s" % (d[:64], m.group(5), m.group(2), m.group(3), m.group(4)))
