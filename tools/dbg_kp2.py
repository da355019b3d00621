import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dbg_kp_lib as d
for rep in range(3):
    for B in (64, 128, 256):
        d.run(B, 23, 2, 24, {})
        d.run(B, 23, 4, 24, {})
        d.run(B, 37, 2, 24, {"AW_PERSISTENT_TILE": "4"})
