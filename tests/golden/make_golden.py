#!/usr/bin/env python
"""Regenerates runtests_known_answers.json from the reference's own test script.

The reference (Julia + ApproxFun) cannot run in the build container, so the only golden values it offers for the
hot path are the known-answer / cross-consistency asserts of test/runtests.jl.  This script reads that file where it
lies (/root/reference/test/runtests.jl -- present in the build container only, never on the GPU box), checks that
every assert the oracle is pinned against is still there, extracts the numeric constants, and writes the JSON that
tests/test_oracle_pins.py and the GPU tests load.  Run:  python tests/golden/make_golden.py"""
import json
import os
import re
import sys

REF = os.environ.get("POLTERGEIST_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    path = os.path.join(REF, "test", "runtests.jl")
    src = open(path, encoding="utf-8").read().splitlines()

    def find(pattern):
        for i, line in enumerate(src, 1):
            if re.search(pattern, line):
                return i, line
        raise SystemExit(f"assert not found in {path}: {pattern}")

    rtol = float(2.220446049250313e-16) ** 0.5          # Julia's default `≈`
    l1, s1 = find(r"@test l_exp ≈ ")
    l2, _ = find(r"@test l_exp2 ≈ ")
    lyap = re.search(r"≈\s*([0-9.]+)", s1).group(1)
    l3, s3 = find(r"@test sigmasq_A ≈ ")
    var = re.search(r"≈\s*([0-9.]+)", s3).group(1)
    l4, _ = find(r"tupling\(-4,0\.\.4\.\)\)\[1:10,1:10\]\) ≈ \(-1/4\)\.\^\(0:9\)")
    l5, _ = find(r"doubling\(PeriodicSegment\(6\.,7\.\)\)\)\[1:10,1:10\]\) ≈ \[1\.;zeros\(9\)\]")
    l6, _ = find(r"@test Transfer\(lan\)\[1:100,1:100\] ≈ L_lan\[1:100,1:100\]")
    l7, s7 = find(r"c\.\^\(1:5\) \+ \(1-c\)\.\^\(1:5\)")
    eig_tol = float(re.search(r"<\s*([0-9.e-]+)\)", s7).group(1))
    l8, s8 = find(r"ρ1f\.\(pts\) - ρ2f\.\(pts\)")
    l8b, _ = find(r"ρ2b\.\(pts\) - ρ2ba\.\(pts\)")
    mult = float(re.search(r"<\s*([0-9]+)eps", s8).group(1))
    l9, _ = find(r"@test M2Fexp ≈ \(Transfer\(M2f\)\*Fun\(exp")
    l10, _ = find(r"@test M2b\.\(test_x\) ≈ test_f")
    l11, _ = find(r"@test all\(diag\(Lfi\[1:10,1:10\]\) \.≈ 1 \./ \(φ")
    l12, _ = find(r"@test lyapunov\(doubleK\) ≈ 2l_exp")
    l13, _ = find(r"lancov\[1\] \+ 2sum\(lancov\[2:end\]\) ≈ birkhoffvar\(lan,A\)")
    out = {
        "_source": "wormell/Poltergeist.jl test/runtests.jl (known-answer asserts; the reference ships no stored arrays). "
                   "Extracted by tests/golden/make_golden.py; the reference itself cannot run here (no Julia).",
        "lanford_lyapunov": {"value": lyap, "cite": f"test/runtests.jl:{l1}-{l2}", "rtol": rtol},
        "lanford_birkhoffvar_x2": {"value": var, "cite": f"test/runtests.jl:{l3}", "rtol": rtol},
        "tupling_m4_diag": {"formula": "(-1/4)^(0:9)", "map": "tupling(-4, 0..4)", "cite": f"test/runtests.jl:{l4}"},
        "doubling_periodic_6_7_diag": {"formula": "[1, 0 x 9]", "map": "doubling(PeriodicSegment(6,7))", "cite": f"test/runtests.jl:{l5}"},
        "modulomap_lanford_block": {"formula": "Transfer(modulomap(5x/2-x^2/2, 0..1))[1:100,1:100] ~ Transfer(lanford())[1:100,1:100]",
                                    "cite": f"test/runtests.jl:{l6}"},
        "sin_asin_top5_eigs": {"formula": "c^k + (1-c)^k, k=1..5, c=1/pi", "atol": eig_tol, "cite": f"test/runtests.jl:{l7}"},
        "acim_cross_consistency": {"formula": "max |rho_Fourier - rho_Chebyshev| at 200 points < 2000 eps; Fwd vs Rev; supplied vs dual derivative",
                                   "atol": mult * 2.220446049250313e-16, "cite": f"test/runtests.jl:{l8}-{l8b}"},
        "transfer_exp": {"formula": "transfer(M2f, exp) ~ Transfer(M2f) * Fun(exp)", "cite": f"test/runtests.jl:{l9}"},
        "newton_round_trip": {"formula": "M2b(mapinv(M2b,1,y)) ~ y on 19 points", "cite": f"test/runtests.jl:{l10}"},
        "induced_golden_diag": {"formula": "1/((phi^k - 1) phi^k), k=1..10", "cite": f"test/runtests.jl:{l11}"},
        "composition_double_lyapunov": {"formula": "lyapunov(lan o lanshift o shiftmap) ~ 2 lyapunov(lan)", "cite": f"test/runtests.jl:{l12}"},
        "covariance_sums": {"formula": "cov[1] + 2 sum(cov[2:end]) ~ birkhoffvar(lan, A)", "cite": f"test/runtests.jl:{l13}"},
    }
    dst = os.path.join(HERE, "runtests_known_answers.json")
    json.dump(out, open(dst, "w"), indent=1, ensure_ascii=False)
    print("wrote", dst)


if __name__ == "__main__":
    sys.exit(main())
