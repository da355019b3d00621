#!/usr/bin/env python
"""Per-source-line warp-stall samples from an .ncu-rep (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py rep.ncu-rep [kernel-substring] [top N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ''
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kernel, fname, hdr = None, None, None
agg = {}
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
            s = int(r[hdr.index('# Samples')] or 0)
            ins = int(r[hdr.index('Instructions Executed')] or 0)
        except (ValueError, IndexError):
            continue
        if s or ins:
            key = (kernel.split('(')[0][-40:], fname, int(r[0]))
            a = agg.setdefault(key, [0, 0, r[1]])
            a[0] += s; a[1] += ins
for k in sorted({k[0] for k in agg}):
    items = [(v[0], v[1], key[1], key[2], v[2]) for key, v in agg.items() if key[0] == k]
    tot = sum(i[0] for i in items) or 1
    toti = sum(i[1] for i in items) or 1
    print(f'== {k}: {tot} samples, {toti} warp-instructions')
    for s, ins, f, ln, text in sorted(items, reverse=True)[:top]:
        print(f'{100.0*s/tot:6.2f}%  inst {100.0*ins/toti:5.2f}%  {f}:{ln}: {text.strip()[:110]}')
    # per source file: the transform code lives in aw_fft*.cuh, the producer / MAC roles in aw_persistent.cu (and the PTX wrappers
    # and cmac2f of aw_fft_blocks.cuh lines < 123).  "FFT share" = warp-instructions issued by the transform code: with the
    # kernel's issue-slot utilisation (smsp__issue_active) it gives the SM throughput the FFT warps reach on their own.
    by_file = {}
    for s_, ins, f, ln, text in items:
        role = 'transforms (aw_fft_reg.cuh, aw_fft.cuh, aw_fft_blocks.cuh lines 123-340)' if (f in ('aw_fft_reg.cuh', 'aw_fft.cuh') or (f == 'aw_fft_blocks.cuh' and 123 <= ln < 341)) else \
               ('frame operand loads, Nyquist sums (aw_fft_blocks.cuh >= line 341)' if (f == 'aw_fft_blocks.cuh' and ln >= 341) else
                'multiply-accumulate + PTX wrappers (aw_fft_blocks.cuh < line 123)' if f == 'aw_fft_blocks.cuh' else f)
        a = by_file.setdefault(role, [0, 0])
        a[0] += s_; a[1] += ins
    for role, (s_, ins) in sorted(by_file.items(), key=lambda kv: -kv[1][1]):
        print(f'   by role: inst {100.0*ins/toti:5.1f}%  samples {100.0*s_/tot:5.1f}%  {role}')
