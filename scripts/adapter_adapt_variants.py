"""ma::adapt through the adapter against ma::adapt with the reference's own size field (tests/adapter/adapter_check.cc:
mag_adapter_adapt_check2) for LogAniso strict / fast and Aniso fast; prints the report rows.  usage: adapter_adapt_variants.py [jitter]"""
import ctypes as C, numpy as np
L=C.CDLL("core_b200/lib/libmag_ma.so")
L.mag_adapter_adapt_check2.argtypes=[C.c_int,C.c_int,C.c_double,C.c_int,C.c_int,C.c_int,C.c_void_p]
L.mag_adapter_set_adapt_jitter.argtypes=[C.c_double]
import sys
L.mag_adapter_set_adapt_jitter(float(sys.argv[1]) if len(sys.argv) > 1 else 0.0)
for log,fp in ((1,0),(1,1),(0,1)):
    out=np.zeros(11); rc=L.mag_adapter_adapt_check2(10,3,1.0,2,log,fp,out.ctypes.data_as(C.c_void_p)); print("log",log,"fp",fp,"rc",rc, out.tolist(), flush=True)
