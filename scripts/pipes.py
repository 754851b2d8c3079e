"""Multiplier-array microbenchmarks (b200zk_microbench): IMAD.WIDE, Montgomery multiplications, FP64 DFMA, and IMAD.WIDE
issued together with DFMA from the same warps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noir_backend_using_gnark_b200 as zk

ctx = zk.Context(0)
names = {0: "IMAD.WIDE.U32 /s", 1: "fp mul /s", 2: "fr mul /s", 3: "DFMA /s", 4: "IMAD.WIDE /s while issuing 2x as many DFMA"}
for w in range(5):
    v = ctx.microbench(w)
    print("%-48s %.3e  (%.1f per clk per SM at 1.965 GHz x 148 SMs)" % (names[w], v, v / 1.965e9 / 148))
