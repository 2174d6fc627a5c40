"""ma::adapt through the adapter vs the reference on a 2-D box (triangles): report rows of mag_adapter_adapt_check2."""
import ctypes as C, numpy as np
L=C.CDLL("core_b200/lib/libmag_ma.so")
L.mag_adapter_adapt_check2.argtypes=[C.c_int,C.c_int,C.c_double,C.c_int,C.c_int,C.c_int,C.c_void_p]
L.mag_adapter_set_adapt_dim.argtypes=[C.c_int]; L.mag_adapter_set_adapt_jitter.argtypes=[C.c_double]
L.mag_adapter_set_adapt_dim(2)
for jit,log,fp in ((0.0,0,0),(0.25,1,1)):
    L.mag_adapter_set_adapt_jitter(jit)
    out=np.zeros(11); rc=L.mag_adapter_adapt_check2(60,3,1.0,2,log,fp,out.ctypes.data_as(C.c_void_p)); print("2d jitter",jit,"log",log,"fp",fp,"rc",rc,out.tolist(),flush=True)
