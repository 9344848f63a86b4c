import ctypes as C, sys
sys.path.insert(0, '.')
from tools import benchlib
L = benchlib.load()
for na in (0, 4, 8):
    v = C.c_double()
    rc = L.ptb_probe_ffma2_issue(0, na, C.byref(v))
    print(na, rc, v.value)
