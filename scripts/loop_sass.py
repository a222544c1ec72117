"""Print the SASS of the Tsit5 step loop of a cached cubin. Usage: loop_sass.py <pattern> [minblocks]"""
import re, subprocess, sys, glob, os
cache = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scimlexpectations.jl_b200", "_cache")
pat = sys.argv[1] if len(sys.argv) > 1 else "du[1] = -u[0]"
mbw = sys.argv[2] if len(sys.argv) > 2 else None
for cu in sorted(glob.glob(cache + "/*.cu")):
    t = open(cu).read()
    if not (pat in t and "#define B200E_FP32 0" in t and "#define B200E_SOLVER 0" in t):
        continue
    mb = re.search(r"#define B200E_MINBLOCKS (\d+)", t).group(1)
    if mbw and mb != mbw: continue
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "b200e_reduce", cu[:-3] + ".cubin"], capture_output=True, text=True).stdout
    lines = [l for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    addr = lambda l: int(re.match(r"\s+/\*([0-9a-f]{4})\*/", l).group(1), 16)
    best = None
    for i, l in enumerate(lines):
        m = re.search(r"BRA\s+(?:[!A-Z0-9]+,\s*)?0x([0-9a-f]+)", l)
        if m and int(m.group(1), 16) < addr(l):
            j0 = next(k for k, x in enumerate(lines) if addr(x) >= int(m.group(1), 16))
            body = lines[j0:i + 1]
            if sum("DFMA" in b for b in body) >= 40 and (best is None or len(body) < len(best)): best = body
    print("#", os.path.basename(cu), "minblocks", mb)
    for b in best:
        print(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", b).strip())
    break
