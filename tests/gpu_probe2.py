import sys, time
import numpy as np
from common import *
from oracle.oracle import Oracle

def pair(name, lib=None, n=400, nint=1, **kw):
    w = World(**SMALL)
    g = LtransLib(lib); o = Oracle()
    prm = make_params(w, n, **kw)
    setup(g, w, prm, n); setup(o, w, prm, n)
    for s in range(1, nint+1):
        g.step(1, s); o.step(1, s)
        fg, fo = g.fetch(), o.fetch()
        dz = np.abs(fg['z']-fo['z'])
        idx = np.argsort(-dz)[:5]
        print(name, 'step', s, 'max dz', dz.max(), 'n>1e-12', (dz>1e-12).sum(), 'top', idx.tolist(), dz[idx].tolist(), flush=True)
    g.destroy(); o.destroy()

passive = dict(HTurbOn=0, VTurbOn=1, Behavior=0, settlementon=0, mortality=0)
pair('fma', None, nint=4, **passive)
pair('nofma', ROOT + '/dbg/libltrans_nofma.so', nint=4, **passive)
