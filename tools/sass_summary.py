#!/usr/bin/env python
"""Static SASS instruction counts per kernel of the in-tree libfairmarl.so (cuobjdump -sass): which kernels carry TMA bulk
copies (UBLKCP), warp reductions (REDUX), float64 arithmetic, local-memory traffic ...  usage: sass_summary.py [lib] > profiles/..."""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "fair-marl_b200/libfairmarl.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
groups = [("UBLKCP (TMA bulk copy)", r"^UBLKCP"), ("REDUX (warp reduce)", r"^REDUX"), ("ATOM/RED (atomics)", r"^(ATOM|RED)\b"), ("MEMBAR", r"^MEMBAR"),
          ("DADD/DMUL/DFMA (fp64)", r"^(DADD|DMUL|DFMA)"), ("MUFU", r"^MUFU"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDG", r"^LDG"), ("STG", r"^STG"),
          ("LDL/STL (local)", r"^(LDL|STL)"), ("SHFL", r"^SHFL"), ("BAR", r"^BAR"), ("ACQBULK/griddepcontrol", r"^(ACQBULK|PREEXIT)")]
counts, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); counts[name] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and name:
        op = m.group(1); counts[name]["instructions"] += 1
        for label, pat in groups:
            if re.match(pat, op): counts[name][label] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"SASS instruction counts per kernel of {lib} (cuobjdump -sass; static counts); cubin architectures: {', '.join(arch)}\n")
for (mangled, c), nice in sorted(zip(counts.items(), names), key=lambda t: t[1]):
    print(nice[:150])
    print("   instructions: %d, " % c["instructions"] + ", ".join(f"{k}: {c[k]}" for k, _ in groups if c[k]))
