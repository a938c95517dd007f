import sys, os
sys.path.insert(0, "/root/repo")
from basic_iterative_solvers_b200 import capi, host
K = 200
with capi.Context(0) as ctx:
    for name, method, pre in (("HPCG-128", "cg", "none"), ("Anderson,Lx=100,Ly=100,Lz=50,ranpot=5.0", "gm", "j")):
        s = host.BenchSession(ctx, name, method, pre, 10)
        s.prepare(10)
        r = s.run(K)
        s.close()
        print(name, method, pre, "res_check_len", os.environ.get("BIS_RES_CHECK_LEN", "1"), f"{r['device_ms'] / K:.4f} ms/iter (device), {r['wall_ms'] / K:.4f} wall", flush=True)
