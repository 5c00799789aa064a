"""Ad-hoc GPU probe (not a pytest file): parity of a few configurations, printed."""
import sys, time
import numpy as np
from common import *
from oracle.oracle import Oracle

def pair(name, n=400, nexternal=4, **kw):
    w = World(**SMALL)
    hab = w.habitat()
    g = LtransLib(); o = Oracle()
    prm = make_params(w, n, **kw)
    setup(g, w, prm, n, habitat=hab); setup(o, w, prm, n, habitat=hab)
    t = time.time(); rg = run(g, w, nexternal); tg = time.time() - t
    t = time.time(); ro = run(o, w, nexternal); to = time.time() - t
    fg, fo = g.fetch(), o.fetch()
    res = compare(fg, fo, w)
    print(name, 'rc', rg, ro, 'gpu_s %.2f cpu_s %.2f' % (tg, to), res, 'stats', g.stats().tolist(), o.stats().tolist(), flush=True)
    eg, eo = g.drain_events(), o.drain_events()
    if eg != eo: print('  EVENTS differ', eg[:5], eo[:5])
    g.destroy(); o.destroy()

if __name__ == '__main__':
    passive = dict(HTurbOn=0, VTurbOn=0, Behavior=0, settlementon=0, mortality=0)
    pair('passive', **passive)
    pair('hturb', **dict(passive, HTurbOn=1))
    pair('vturb', **dict(passive, VTurbOn=1))
    pair('salttemp', **dict(passive, SaltTempOn=1))
    for b in range(1, 8):
        pair('behav%d' % b, Behavior=b, settlementon=0, mortality=0, HTurbOn=0, VTurbOn=0)
    pair('full4', Behavior=4, pediage=3600.0, deadage=9000.0, swimstart=0.0)
