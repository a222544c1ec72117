"""Generates the committed fixtures under tests/golden/.

The reference is Julia and cannot run here, so the fixtures are (a) the numbers the reference itself
publishes (README.md:50-55, :65-69, :73-78 -- copied as data into readme_linear2.json) and (b) outputs
of the CPU oracle on fixed seeded inputs, after the oracle has reproduced (a) (tests/test_oracle_golden.py
pins it to the README digits).  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from tests.sampling_util import sample_points  # noqa: E402

README = {
    "source": "reference README.md:50-55 (analytical), :65-69 (Koopman u), :73-78 (resid)",
    "analytical": [1.5433991194037804, -1.120209038276938],
    "koopman_u": [1.5433860531082695, -1.1201922503747408],
    "koopman_resid": [7.193424502016654e-5, 5.2074632876847327e-5],
    "quadalg": "CubatureJLh", "ireltol": 1e-3, "iabstol": 1e-3,
}


def main():
    with open(os.path.join(HERE, "readme_linear2.json"), "w") as f:
        json.dump(README, f, indent=1)
    out = {}
    for name in ("linear2", "lotka_volterra", "lorenz63", "decay1", "oscillator_noise", "gbm4", "gbm4_lamba"):
        n = 256 if name not in ("oscillator_noise", "gbm4", "gbm4_lamba") else 64
        x = sample_points(name, n, seed=11)
        if name == "gbm4_lamba":  # test/processnoise.jl:15: LambaEM(), abstol = reltol = 1e-3
            orc = oracle.OracleProblem.builtin("gbm4", solver=oracle.SOLVER_LAMBA_EM, abstol=1e-3, reltol=1e-3)
        else:
            orc = oracle.OracleProblem.builtin(name)
        dx, c, steps = orc.eval_nodes(x, weight_by_pdf=True, want_steps=True)
        s, s2, c2 = orc.reduce_samples(x, weight_by_pdf=False)
        out[f"{name}_x"] = x
        out[f"{name}_dx"] = dx
        out[f"{name}_steps"] = steps
        out[f"{name}_sum"] = s
        out[f"{name}_sumsq"] = s2
        out[f"{name}_counters"] = np.array([c["nf"], c["naccept"], c["nreject"], c["nfail"]], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
