"""Static instruction mix of the Tsit5 step loop of a cached cubin (innermost backward branch after a VOTE)."""
import re, subprocess, sys, glob, os, collections
cache = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scimlexpectations.jl_b200", "_cache")
pat = sys.argv[1] if len(sys.argv) > 1 else "du[1] = -u[0]"
fp = sys.argv[2] if len(sys.argv) > 2 else "0"
for cu in sorted(glob.glob(cache + "/*.cu")):
    t = open(cu).read()
    if pat in t and f"#define B200E_FP32 {fp}" in t and "#define B200E_SOLVER 0" in t:
        cubin = cu[:-3] + ".cubin"
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", "b200e_reduce", cubin], capture_output=True, text=True).stdout
        lines = [l for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
        addr = lambda l: int(re.match(r"\s+/\*([0-9a-f]{4})\*/", l).group(1), 16)
        # backward branches: (from, to); pick the one whose body contains the most DFMAs
        best = None
        for i, l in enumerate(lines):
            m = re.search(r"BRA\s+(?:[!A-Z0-9]+,\s*)?0x([0-9a-f]+)", l)
            if m and int(m.group(1), 16) < addr(l):
                j0 = next(k for k, x in enumerate(lines) if addr(x) >= int(m.group(1), 16))
                body = lines[j0:i + 1]
                n = sum("DFMA" in b for b in body)
                if best is None or (n >= 60 and len(body) < len(best)) or (sum("DFMA" in b for b in best) < 60 and n > sum("DFMA" in b for b in best)):
                    best = body
        ops = collections.Counter()
        for b in best:
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", b)
            ops[m.group(1)] += 1
        dp = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
        print(os.path.basename(cubin), "loop instrs", len(best), "DP", dp, "non-DP", len(best) - dp)
        print("  ", dict(ops.most_common(18)))
