#!/usr/bin/env python
"""What bounds the lag kernel's inner loop?  DFMA rates of register-only probes on the current GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acfcalculator_b200 as A
peak, mhz = A.measure_dfma_peak()
print("pure DFMA chains (2 reusable operands), 4 warps/SMSP: %.2f TFLOP/s  (nominal 37.22)" % peak)
LABELS = {0: "no loads (+1 DMUL)", 1: "1 LDS.128 distinct / 18 DFMA", 2: "1 LDS.128 broadcast / 18 DFMA",
          3: "1 LDS.64 distinct / 18 DFMA (+1 DMUL)", 4: "2 LDS.64 distinct / 18 DFMA",
          5: "1 LDS.128 distinct, result used later (+1 DMUL +1 DADD)", 6: "2 SHFL / 18 DFMA",
          7: "pure DADD chains (slot rate as DFMA-equivalent)", 8: "DADD->DFMA pairs, no loads (slot rate)"}
for wps in (1, 4):
    for kind in range(9):
        tf = A.measure_dfma_probe(10 * wps + kind)
        print("%d warp(s)/SMSP  %-55s %.2f TFLOP/s  (%.3f of pure)" % (wps, LABELS[kind], tf, tf / peak))
