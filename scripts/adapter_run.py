import ctypes as C, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "core_b200/lib/libmag_ma.so"))
L.mag_adapter_check.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
for log_interp, fp in ((0, 0), (0, 1), (1, 0), (1, 1)):
    rep = np.zeros(20)
    rc = L.mag_adapter_check(int(sys.argv[1]) if len(sys.argv) > 1 else 16, log_interp, fp, 0.25, rep.ctypes.data_as(C.c_void_p))
    print("log=%d fp=%d rc=%d" % (log_interp, fp, rc), rep.tolist(), flush=True)
