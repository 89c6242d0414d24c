"""Runs the FP64 roofline-denominator microbenchmark (idp_measure_fp64_tflops) alone, for an ncu capture of its pipe utilisation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from idp_b200 import ContactContext
c = ContactContext(0)
print("fp64 TFLOP/s", c.fp64_tflops())
