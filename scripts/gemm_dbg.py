import os, subprocess, sys
for flags in (0, 1, 2, 3):
    env = dict(os.environ, MSF_TC_DEBUG=str(flags))
    out = subprocess.run([sys.executable, "scripts/gemm_graph_bench.py"], env=env, capture_output=True, text=True).stdout
    print("MSF_TC_DEBUG=%d" % flags)
    for l in out.splitlines():
        if "n= 256 k=   64" in l or "m=37888" in l or "m=18944" in l or "m=  128 n=  32" in l:
            print("  ", l)
