import sys, os, ctypes as C
import numpy as np
os.environ['ORA_TRACE_ID'] = '1761'
from common import *
from oracle.oracle import Oracle
PASSIVE = dict(HTurbOn=0, VTurbOn=1, Behavior=0, settlementon=0, mortality=0)
w = World(**SMALL); n = 2000
g = LtransLib(ROOT + '/dbg/libltrans_dbg.so'); o = Oracle()
prm = make_params(w, n, **PASSIVE)
setup(g, w, prm, n); setup(o, w, prm, n)
g.lib.ltgpu_debug_trace(g.ctx, C.c_int64(1761), None, 0)
g.step(1, 1); o.step(1, 1)
buf = np.zeros(240)
g.lib.ltgpu_debug_trace(g.ctx, C.c_int64(1761), buf.ctypes.data_as(C.c_void_p), 240)
for i in range(60): print('GPUTRACE', i, *['%.17g' % v for v in buf[4*i:4*i+4]])
