"""cuobjdump -sass of the two shard-sized tree kernels -> profiles/r2_sass_excerpt_select_f_backprop_f.txt: instruction count, opcode
histogram and one example of every memory / special instruction form.   python tools/sass_excerpt.py > profiles/<name>.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "alphazero-al_b200", "libazb200.so")
KERNELS = [("k_select_f_r<C4, VL, AUX, RO> (the shard-sized select of the bench workload; root scored once per launch)",
            "_ZN2az12k_select_f_rINS_2C4ELb1ELb1ELb1ELb0EEEvNS_3DevE16az_search_configiPK7az_rootP7az_leaf"),
           ("k_backprop_f_r<C4, VL, RO> (the shard-sized back-prop)",
            "_ZN2az14k_backprop_f_rINS_2C4ELb1ELb1ELb0EEEvNS_3DevE16az_search_configiiiiPKfS5_S5_S5_S5_PKhPKi")]
SPECIAL = re.compile(r"^(LDG|STG|LDGSTS|LDS|STS|ATOM|RED|MUFU|SHFL|ACQBULK|PREEXIT|DEPBAR|LDGDEPBAR|CCTL|BAR|WARPSYNC|VOTE|REDUX|CALL|MATCH|ELECT|UBLKCP|LDL|STL)")
print("cuobjdump -sass of alphazero-al_b200/libazb200.so (sm_100a), round 2 (final sources): instruction mix and one example of every memory / special instruction form")
print("(LDGSTS = cp.async, LDG/STG.E.ENL2.256 = 256-bit global accesses, ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents, MUFU.RCP = the seed of the branch-free\n"
      "divisions; the count includes the out-of-line rare paths that ptxas places in the kernel's section: plain-operator scoring, logf, statistics, Dirichlet noise)\n")
for title, sym in KERNELS:
    out = subprocess.run(["cuobjdump", "-sass", "-fun", sym, SO], capture_output=True, text=True).stdout
    ins = [l for l in out.splitlines() if re.match(r"\s*/\*[0-9a-f]{4}\*/", l)]
    ops = []
    for l in ins:
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        ops.append(m.group(1) if m else "?")
    hist = collections.Counter(ops)
    print(f"== {title}: {len(ins)} SASS instructions ({len(ins) * 16 / 1024:.1f} KB)")
    print("opcode histogram (top 40): " + ", ".join(f"{k} {v}" for k, v in hist.most_common(40)))
    sp = {k: v for k, v in hist.items() if SPECIAL.match(k)}
    print("memory / special instructions: " + ", ".join(f"{k} x{v}" for k, v in sorted(sp.items())))
    seen = set()
    for l, o in zip(ins, ops):
        if SPECIAL.match(o) and o not in seen:
            seen.add(o)
            print("   " + l.strip()[:150])
    print()
