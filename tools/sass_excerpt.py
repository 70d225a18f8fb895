"""profiles/r02_sass_excerpts.md: opcode census and an excerpt of k_ppe_stream, and the TMA / mbarrier / DSMEM / REDUX counts of the
other pressure kernels, from cuobjdump -sass of the built library.   python tools/sass_excerpt.py > profiles/r02_sass_excerpts.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "computational-fluid-dynamics_b200", "lib", "libpm.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s*/\*[0-9a-f]{4,5}\*/", line):
        funcs[cur].append(re.sub(r"\s*/\*[0-9a-f]{16}\*/\s*$", "", line.rstrip()))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
name = next(n for n in funcs if "k_ppe_stream" in n and "ILi0ELi0E" in n)
lines = funcs[name]
op = lambda l: re.sub(r"^\s*/\*[0-9a-f]+\*/\s+(@!?U?P\w+\s+)?", "", l).split()[0].split(".")[0]
ops = collections.Counter(op(l) for l in lines)
print("# r02 SASS evidence (cuobjdump -sass lib/libpm.so, arch sm_100a; tools/sass_excerpt.py)\n")
print(f"## {demangle(name)} (cavity form): {len(lines)} instructions; two unrolled 8-tick bodies (steady / chunk ends)\n")
print("| opcode | count |\n|---|---|")
for k, v in ops.most_common(20):
    print(f"| {k} | {v} |")
print("\nNo BAR (no block barrier), no UTMALDG (the rings are filled by LDGSTS = cp.async), no local memory in the loop:\n")
for pat in ("LDGSTS", "LDGDEPBAR", "DEPBAR.LE", "SHFL", "LDS.128", "STS.128", "STG.E.128", "DSETP", "LDL", "STL", "BAR.SYNC"):
    hits = [l for l in lines if pat in l]
    ex = re.sub(r"/\*.*?\*/", "", hits[0]).strip() if hits else ""
    print(f"* `{pat}` x {len(hits)}" + (f":  `{ex}`" if ex else ""))
i = next(i for i, l in enumerate(lines) if "DEPBAR.LE" in l)
print("\nFirst instructions of a tick (wait for the row, take it from the ring, refill the slot, first half-sweeps):\n\n```")
print("\n".join(lines[i:i + 56]))
print("```\n\n## TMA / mbarrier / distributed shared memory / REDUX in the other pressure kernels (instruction counts)\n")
print("| kernel | UTMALDG | SYNCS (mbarrier) | STAS (st.async) | REDUX |\n|---|---|---|---|---|")
for n, ls in funcs.items():
    d = demangle(n)
    if not any(k in d for k in ("k_ppe_tiled<Fast, 0, 1, 4, 0>", "k_ppe_tiled<Fast, 1, 1, 4, 0>", "k_ppe_tiled<Exact, 0, 1, 3, 0>", "k_ppe_cluster<Fast, 0, false>", "k_ppe_cluster<Fast, 1, true>")):
        continue
    c = lambda p: sum(1 for l in ls if p in l)
    print(f"| `{d}` | {c('UTMALDG')} | {c('SYNCS')} | {c('STAS')} | {c('REDUX')} |")
