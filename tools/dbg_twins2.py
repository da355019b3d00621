import os, sys
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dbg_twins.py")).read().split("\nrun(1024, 512, 65536, 132, 4096)")[0]
exec(src)
run(1024, 512, 4320, 40, 4096)
run(1024, 512, 65536, 40, 4096)
run(2048, 64, 4320, 80, 256)
run(2048, 128, 4320, 60, 512)
run(4096, 256, 4320, 40, 1024)
run(2048, 1024, 4320, 12, 4096)
run(2048, 2048, 4320, 8, 4096)
run(37, 512, 4320, 40, 4096, S=2)
