"""Differential run on a B200: random models (tests/fuzz_util.py) x random batch shapes, kernel against oracle.
usage: fuzz_parity.py [first_seed] [count] [ode | large | noise | logpoly]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scimlexpectations_jl_b200 as sme
from tests.fuzz_util import compare, oracle_of, random_model, random_noise_model, with_logpoly_factors

first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
count = int(sys.argv[2]) if len(sys.argv) > 2 else 16
family = sys.argv[3] if len(sys.argv) > 3 else "ode"
bad = 0
for seed in range(first, first + count):
    spec, draw = random_noise_model(seed) if family == "noise" else random_model(seed, nmax=16 if family == "large" else 5)
    if family == "logpoly":
        spec, draw = with_logpoly_factors(seed, spec, draw)
        if not any(int(d[0]) == 5 for d in spec.pdf):
            print(f"# seed {seed}: no positive-support dimension, skipped", flush=True)
            continue
    rng = np.random.default_rng([7, seed])
    orc = oracle_of(spec)
    t0 = time.time()
    try:
        with sme.Handle(spec.replace(chunk_points=int(rng.choice([0, 777, 4096])))) as h:
            big = [1000, 3001] if family == "noise" else [5000, 20011, 65537]
            for n in (1, int(rng.integers(2, 300)), int(rng.choice(big))):
                r = compare(h, spec, orc, draw(n, n), int(rng.integers(0, 2)), bool(rng.integers(0, 2)))
                print(json.dumps({"family": family, "solver": spec.solver, "n_kkl": spec.n_kkl, "seed": seed, "nstate": spec.nstate, "nx": spec.nx, "nout": spec.nout, "nsave": len(spec.save_times),
                                  "schedule": spec.schedule, "reltol": spec.reltol, **r}), flush=True)
    except AssertionError as e:
        bad += 1
        print("PARITY VIOLATION seed", seed, str(e)[:400], flush=True)
    print(f"# seed {seed}: {time.time() - t0:.1f} s", flush=True)
print("violations:", bad)
sys.exit(1 if bad else 0)
