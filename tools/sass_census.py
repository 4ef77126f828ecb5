"""SASS opcode census of the shipped library: which Blackwell-native instructions each kernel uses (no GPU needed).
Usage: python tools/sass_census.py [ast_b200/libast_b200.so] > profiles/rNN_sass_opcodes.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "ast_b200/libast_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCATOMSWS", "HMMA", "SYNCS", "STAS", "REDAS", "ATOMS",
        "RED", "LDS", "STS", "LDG", "STG", "LD.E", "MUFU", "DFMA", "DADD", "BAR", "CCTL", "ERRBAR", "MEMBAR", "UCGABAR"]
per = collections.OrderedDict(); cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); per[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                per[cur][k] += 1
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print(f"{lib}: {len(per)} kernels, {tot['_total']} SASS instructions (sm_100a)")
print("whole library: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]))
print()
short = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0][:70]
print(f"{'kernel':70s} {'instrs':>7s}  tensor / TMA / TMEM / cluster opcodes")
for name, c in sorted(per.items(), key=lambda kv: -kv[1]["_total"]):
    hot = [k for k in ("UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "STTM", "HMMA", "SYNCS", "STAS", "REDAS", "UCGABAR", "DFMA") if c[k]]
    print(f"{short(name):70s} {c['_total']:7d}  " + ", ".join(f"{k} {c[k]}" for k in hot))
