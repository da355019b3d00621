"""GPU stress (not part of the test-suite): long runs of the full-size twin check — every stream must stay bit-identical to its
twin over hundreds of blocks.  usage: python tools/stress_twins.py"""
import os
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dbg_twins.py")).read().split("\nrun(1024, 512, 65536, 132, 4096)")[0]
exec(src)
run(4096, 256, 4320, 300, 1024)
run(2048, 512, 4320, 200, 2048)
run(2048, 64, 4320, 600, 256)
run(2048, 128, 4320, 400, 512)
run(2048, 1024, 4320, 60, 4096)
run(2048, 2048, 4320, 30, 4096)
run(1024, 512, 65536, 140, 4096)
