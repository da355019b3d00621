#!/usr/bin/env python
"""CPU-side: turns the files tools/profile_round.sh left in gpurun_out/ into the committed evidence under profiles/.
usage: python tools/collect_profiles.py r01"""
import collections, csv, io, json, os, re, shutil, subprocess, sys

R = sys.argv[1] if len(sys.argv) > 1 else "r02"
REP = os.environ.get("AW_REP_DIR", "gpurun_out")
os.makedirs("profiles", exist_ok=True)
if os.path.exists(f"gpurun_out/{R}_bench_all.txt"):
    shutil.copy(f"gpurun_out/{R}_bench_all.txt", f"profiles/{R}_bench_all.txt")
for name, skip in [("C2", 100), ("C4", 190)]:
    shutil.copy(f"gpurun_out/{R}_launches_{name}.csv", f"profiles/{R}_launches_{name}.csv")
    rows = [r for r in csv.reader(open(f"profiles/{R}_launches_{name}.csv")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1000.0 if r[ui] in ("nsecond", "ns") else v * 1000.0 if r[ui] in ("msecond", "ms") else v
        agg.setdefault(r[ki], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(f"profiles/{R}_launches_{name}_summary.txt", "w") as f:
        cmd = "python bench.py --steps 100 --warmup 40 --no-cpu --e2e-steps 3 --no-single-block" + ("" if name == "C2" else " --workload C4")
        f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none -s {skip} -c 40 {cmd}\n")
        f.write("(40 consecutive launches inside the timed region; per-launch times under ncu are cold-cache and serialised: compare shares)\n")
        for k, v in agg.items():
            f.write(f"{k[:90]:90s} launches {len(v):3d}  mean {sum(v)/len(v):9.2f} us  share {100*sum(v)/tot:5.1f}%\n")
workloads = ["C2", "C2k1", "C3", "C4", "C5-512", "C5-64", "C5-2048", "C4eq", "C5-4096_k2", "C5-4096_k4"]
with open(f"profiles/{R}_ncu_summary.txt", "w") as f:
    f.write("# ncu summaries (tools/ncu_summary.py + tools/ncu_lines.py over the .ncu-rep files of tools/profile_round.sh)\n"
            "# Captures: ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 45 -c 1 python bench.py --workload <W> --steps 4 --warmup 41 --no-cpu --e2e-steps 3, one B200.\n"
            "# A launch = one 1024-frame call (C2, C4: 4 blocks; C5-64: 16; C5-512, C3: 2; C5-2048: 1); C2k1 = C2 with one block per launch.\n"
            "# Per-launch times under ncu are cold-cache and serialised (compare shares, not absolutes); bench.py numbers are never taken under ncu.\n")
    for w in workloads:
        rep = f"{REP}/{R}_full_{w}.ncu-rep"
        f.write(f"\n##### workload {w}\n")
        f.write(subprocess.run([sys.executable, "tools/ncu_summary.py", rep], capture_output=True, text=True).stdout)
        f.write("--- top source lines by warp-stall samples\n")
        f.write(subprocess.run([sys.executable, "tools/ncu_lines.py", rep, "k_", "10"], capture_output=True, text=True).stdout)
out = {}
for w, kern in [("C2", "k_persistent<8,4>"), ("C2k1", "k_persistent<8,4>"), ("C3", "k_persistent<9,4>"), ("C4", "k_persistent<8,4>"), ("C5-512", "k_persistent<9,4>"), ("C5-64", "k_persistent<6,4>"), ("C5-2048", "k_persistent<11,2>"),
                ("C5-4096", "k_input_rfft<12>")]:
    raw = subprocess.run(["ncu", "-i", f"{REP}/{R}_full_{w if w != 'C5-4096' else 'C5-4096_k2'}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, r = rows[0], rows[1], rows[2]
    def val(name):
        i = h.index(name)
        return float(r[i].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[i]]
    out[w] = {"kernel": kern, "source": f"profiles/{R}_ncu_summary.txt (ncu --set full, one launch, workload {w}; ncu flushes caches around the launch, so write-backs still in L2 at its end are not counted)",
              "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum")}
json.dump(out, open(f"profiles/{R}_traffic.json", "w"), indent=1)
print({k: (round(v["dram_bytes_read"] / 1e6), round(v["dram_bytes_write"] / 1e6)) for k, v in out.items()})

# SASS evidence of the shipped library
sass = subprocess.run(["cuobjdump", "-sass", "airwave_b200/lib/libairwave_cuda.so"], capture_output=True, text=True).stdout
lines_out = ["# SASS evidence (cuobjdump -sass airwave_b200/lib/libairwave_cuda.so), sm_100a — instruction counts per kernel",
             "# tensor-map TMA loads = UTMALDG, TMA-class 1-D bulk copies = UBLKCP (with L2 cache hints), mbarrier = SYNCS.*, cp.async = LDGSTS, named barriers = BAR.SYNC/BAR.ARV,",
             "# programmatic dependent launch = ACQBULK/PREEXIT-class griddepcontrol instructions, float64 = DADD/DMUL",
             "# No tensor-core instruction anywhere (the per-bin contraction has two outputs): no UTCMMA/HMMA/QGMMA.", ""]
cur, counts = None, collections.OrderedDict()
want = re.compile(r"^(UTMALDG|UBLKCP|SYNCS|LDGSTS|BAR\.|FFMA$|FFMA\.|LDS|STS|LDG|STG|DADD|DMUL|DFMA|SHFL|WARPSYNC|HMMA|UTCMMA|QGMMA|MEMBAR|FENCE|ATOMS|REDS|RED\.|ACQBULK|PREEXIT|GRIDDEP)")
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur and want.match(m.group(1)):
        op = m.group(1)
        key = op if op.startswith(("UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "BAR", "FENCE", "MEMBAR", "ACQBULK", "PREEXIT", "GRIDDEP")) else op.split(".")[0]
        counts[cur][key] += 1
lines_out.append(f"# architectures in the fatbin: {sorted(set(re.findall(r'arch = (sm_[0-9a-z]+)', sass)))}")
for k, c in counts.items():
    if c and any(t in k for t in ("k_persistent", "k_fused", "k_fdl_cmac", "k_input_rfft", "k_irfft_out", "k_eq", "k_bank_build")):
        lines_out.append(k)
        lines_out.append("    " + ", ".join(f"{a}:{b}" for a, b in sorted(c.items())))
open(f"profiles/{R}_sass_evidence.txt", "w").write("\n".join(lines_out) + "\n")
