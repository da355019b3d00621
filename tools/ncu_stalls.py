#!/usr/bin/env python
"""Stall reasons per role (transforms / multiply-accumulate / producer+control) from an .ncu-rep with source info.
usage: python tools/ncu_stalls.py rep.ncu-rep [kernel-substring]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else 'k_'
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kernel = fname = hdr = None
roles = {}
def role_of(f, ln):
    if f in ('aw_fft_reg.cuh', 'aw_fft.cuh') or (f == 'aw_fft_blocks.cuh' and 123 <= ln < 341):
        return 'transforms'
    if f == 'aw_fft_blocks.cuh' and ln >= 341:
        return 'frame-operand-loads'
    if f == 'aw_fft_blocks.cuh':
        return 'mac+ptx-wrappers'
    if f.startswith('sm_'):
        return 'intrinsics(' + f + ')'
    return f
for r in rows:
    if not r:
        continue
    if r[0] == 'Function Name':
        kernel = r[1]; continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        hdr = r; continue
    if hdr and len(r) == len(hdr) and kernel and want in kernel and r[0] != '':
        try:
            ln = int(r[0])
        except ValueError:
            continue
        ro = roles.setdefault(role_of(fname, ln), {})
        for i, h in enumerate(hdr):
            if h.startswith('stall_') and '(Not Issued)' not in h or h in ('# Samples', 'Instructions Executed'):
                try:
                    ro[h] = ro.get(h, 0) + int(r[i] or 0)
                except ValueError:
                    pass
for name, ro in sorted(roles.items(), key=lambda kv: -kv[1].get('# Samples', 0)):
    tot = ro.get('# Samples', 0) or 1
    top = sorted(((v, k) for k, v in ro.items() if k.startswith('stall_')), reverse=True)[:7]
    print(f"{name:28s} samples {tot:7d} inst {ro.get('Instructions Executed', 0):10d}  " + "  ".join(f"{k[6:]} {100.0 * v / tot:.0f}%" for v, k in top if v))
