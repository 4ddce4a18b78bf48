"""Dev tool: per kernel of libvcprep.so, the SASS instruction count and the Blackwell / Hopper-class mnemonics it uses
(`cuobjdump -sass`).  usage: sass_mnemonics.py > profiles/rNN_sass_mnemonics.md"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "vision_compression_project_b200/libvcprep.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
NOTE = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "VABSDIFF4", "VIMNMX", "VIADD", "IDP", "ATOMS", "REDUX", "MATCH", "ELECT", "BAR", "UTCMMA", "HMMA"]
kern, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1)); kern[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        kern[cur]["_n"] += 1
        op = m.group(1)
        for k in NOTE:
            if op == k or op.startswith(k + "_") or (k in ("VIMNMX", "VIADD", "IDP", "BAR") and op == k):
                kern[cur][k] += 1
print("# SASS of libvcprep.so (sm_100a, `cuobjdump -sass`, `tools/sass_mnemonics.py`): instruction count and the Blackwell / Hopper-class mnemonics per kernel\n")
print("`UBLKCP` = `cp.async.bulk` (TMA 1-D bulk copy, global -> shared), `SYNCS` = mbarrier arrive / expect_tx / try_wait, `UCGABAR_*` = cluster barrier,")
print("`VABSDIFF4` / `VIMNMX` / `VIADD` (incl. `.16x2`) / `IDP.4A` = packed byte / half-word integer SIMD, `ATOMS` = shared-memory atomics, `REDUX` / `MATCH` = warp reductions.")
print("Full listing of the dominant kernel: `r02_k_lz_CfgDefault.sass`.  No tensor-core (`UTC*MMA`, `HMMA`) instruction anywhere: nothing on this path is a contraction.\n")
print("| kernel | SASS instructions | mnemonics of note |\n|---|---|---|")
for name, c in sorted(kern.items(), key=lambda kv: -kv[1]["_n"]):
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    short = re.sub(r"\(anonymous namespace\)::", "", short)
    notes = ", ".join(f"{k} x{c[k]}" for k in sorted(c) if k != "_n") or "-"
    print(f"| `{short}` | {c['_n']} | {notes} |")
