#!/usr/bin/env python3
"""Print the field-product microbenchmarks (vrfs_measure_mac32_peak variants 2, 3) of the library selected by VRFS_B200_LIB."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ark_ec_vrfs_b200 as vrfs
e = vrfs.Engine(0)
for v, name in ((2, "montmul_chain1"), (3, "montmul_chain2")):
    macs, mhz = e.measure_mac32_peak(v)
    print("%s: %.2f G products/s" % (name, macs / 136 / 1e9))
