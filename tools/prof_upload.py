import sys, os, time, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np
import csp_b200 as cb
from csp_b200 import device as D, _native as N
root, obj = cb.synthetic.config_tree("C2")
flat = root.tree.flat()
L = N.dev()
for it in range(6):
    t0 = time.perf_counter()
    d, keep = D.make_desc(flat)
    t1 = time.perf_counter()
    h = C.c_void_p()
    N.dchk(L.csp_tree_upload(C.byref(d), C.byref(h)))
    t2 = time.perf_counter()
    L.csp_tree_free(h)
    t3 = time.perf_counter()
    print("make_desc %.2f ms  upload %.2f ms  free %.2f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3))
