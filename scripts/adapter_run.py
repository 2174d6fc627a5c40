"""The C++ adapter against the unmodified reference in one process on the GPU box (libmag_ma.so is prebuilt where the
reference tree exists): parity report and wall-clock of the five whole-mesh sweeps (split, collapse, bad quality, min quality,
max length) done by the reference alone, by the adapter's bulk entry points (MDS export + upload + device sweep + flag
write-back, five times) and by the unmodified reference loops with the adapter plugged into ma::Input.
usage: adapter_run.py [cells per side, default 16] [host threads of the adapter's MDS walk, default 1] [how many of the 4 (interpolation, arithmetic) cases, default 4]"""
import ctypes as C, numpy as np, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = C.CDLL(os.path.join(ROOT, "core_b200/lib/libmag_ma.so"))
L.mag_adapter_check.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
L.mag_adapter_times.argtypes = [C.c_void_p]
L.mag_adapter_times2.argtypes = [C.c_void_p]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
L.mag_adapter_set_threads.argtypes = [C.c_int]
L.mag_adapter_set_threads(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ncfg = int(sys.argv[3]) if len(sys.argv) > 3 else 4
for log_interp, fp in ((0, 0), (0, 1), (1, 0), (1, 1))[:ncfg]:
    rep, t = np.zeros(20), np.zeros(3)
    rc = L.mag_adapter_check(n, log_interp, fp, 0.25, rep.ctypes.data_as(C.c_void_p))
    L.mag_adapter_times(t.ctypes.data_as(C.c_void_p))
    t2 = np.zeros(15)
    L.mag_adapter_times2(t2.ctypes.data_as(C.c_void_p))
    print("n=%d log=%d fp=%d rc=%d counts=%s  reference %.3f s | adapter bulk %.3f s first use (%.1fx), %.3f s warm re-export (%.1fx) | reference loops over the adapter %.3f s (%.1fx)"
          % (n, log_interp, fp, rc, rep[:3].astype(np.int64).tolist(), t[0], t[1], t[0] / t[1], t2[0], t[0] / t2[0], t[2], t[0] / t[2]), flush=True)
    names = ("export", "revalidate", "upload", "flags_in", "device", "flags_out", "refresh")
    for r, label in ((0, "first"), (1, "warm ")):
        print("    %s ms: " % label + "  ".join("%s %.1f" % (k, 1e3 * v) for k, v in zip(names, t2[1 + 7 * r: 8 + 7 * r])), flush=True)
