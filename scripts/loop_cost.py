"""Issue-cost model of the Tsit5 step loop of a cached cubin (see profiles/r1_dp_operands_ubench.txt):
a DP instruction costs 2 clk with <= 2 distinct 64-bit register sources and 3 clk with 3; IMAD ~0.7 clk;
ALU / LDS / LDC are hidden behind the FP64 pipe.  Usage: loop_cost.py <pattern in model source> [fp32]"""
import re, subprocess, sys, glob, os, collections
cache = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scimlexpectations.jl_b200", "_cache")
pat = sys.argv[1] if len(sys.argv) > 1 else "du[1] = -u[0]"
show = len(sys.argv) > 2 and sys.argv[2] == "show"
for cu in sorted(glob.glob(cache + "/*.cu")):
    t = open(cu).read()
    if not (pat in t and "#define B200E_FP32 0" in t and ("#define B200E_SOLVER 0" in t or "#define B200E_SOLVER 2" in t)):
        continue
    mb = re.search(r"#define B200E_MINBLOCKS (\d+)", t).group(1)
    cubin = cu[:-3] + ".cubin"
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "b200e_reduce", cubin], capture_output=True, text=True).stdout
    lines = [l for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    addr = lambda l: int(re.match(r"\s+/\*([0-9a-f]{4})\*/", l).group(1), 16)
    best = None
    for i, l in enumerate(lines):
        m = re.search(r"BRA\s+(?:[!A-Z0-9]+,\s*)?0x([0-9a-f]+)", l)
        if m and int(m.group(1), 16) < addr(l):
            j0 = next(k for k, x in enumerate(lines) if addr(x) >= int(m.group(1), 16))
            body = lines[j0:i + 1]
            n = sum("DFMA" in b for b in body)
            if n >= 40 and (best is None or len(body) < len(best)):
                best = body
    ops = collections.Counter(); dp2 = dp3 = 0; three = []
    for b in best:
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)(?:\.[A-Z0-9_.]+)?\s+(.*?);", b)
        op, args = m.group(1), m.group(2)
        ops[op] += 1
        if op in ("DFMA", "DMUL", "DADD", "DSETP"):
            srcs = [a.strip() for a in args.split(",")]
            srcs = srcs[1:] if op != "DSETP" else srcs[2:]
            regs = set(re.sub(r"[-|]|\.reuse", "", s) for s in srcs if re.match(r"^[-|]*R\d+", s.strip()))
            regs.discard("RZ")
            if len(regs) >= 3: dp3 += 1; three.append(b.split("*/")[1].strip())
            else: dp2 += 1
    imad = ops["IMAD"]
    cost = 2 * dp2 + 3 * dp3 + 0.7 * imad
    print(f"{os.path.basename(cubin)} minblocks {mb}: loop {len(best)} instrs, DP {dp2 + dp3} (3-reg: {dp3}), IMAD {imad}, model cost {cost:.0f} clk")
    if show:
        for s in three: print("    3reg:", s)
