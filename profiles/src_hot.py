"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line."""
import csv, collections, sys
path = sys.argv[1]; root = sys.argv[2] if len(sys.argv) > 2 else '/root/repo/ltransv.2b_b200/csrc/'
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rd = csv.reader(open(path)); cur_file = cur_fn = hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in rd:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': cur_fn = r[1][:24]; continue
    if r[0] == 'Line No': hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < 12 or r[2] != '-': continue
    try:
        line = int(r[0]); ie = int(r[ix['Instructions Executed']] or 0); te = int(r[ix['Thread Instructions Executed']] or 0); s = int(r[ix['# Samples']] or 0)
    except Exception: continue
    a = agg[(cur_fn, cur_file, line)]; a[0] += ie; a[1] += te; a[2] += s
src = {}
import os
for fn in os.listdir(root):
    if fn.endswith(('.cuh', '.cu', '.h')): src[fn] = open(root + fn).read().split('\n')
for kern in sorted(set(k[0] for k in agg)):
    items = [(k, v) for k, v in agg.items() if k[0] == kern]
    tot = sum(v[0] for k, v in items)
    print('=====', kern, 'warp inst', tot, 'thread inst', sum(v[1] for k, v in items), 'samples', sum(v[2] for k, v in items))
    items.sort(key=lambda kv: -kv[1][0])
    for (fn, fl, ln), v in items[:top]:
        text = src[fl][ln - 1].strip()[:88] if fl in src and ln - 1 < len(src[fl]) else ''
        print(f'{100*v[0]/max(1,tot):5.1f}% act {v[1]/max(1,v[0]):4.1f} smp {v[2]:6d} {fl}:{ln} {text}')
