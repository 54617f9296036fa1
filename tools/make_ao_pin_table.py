"""Independent pin for the AO evaluator: values and gradients of one s and one p contracted Gaussian shell at
ten points, computed here with mpmath at 50 digits DIRECTLY from the textbook formulas -- nothing from this
repository is imported (no molgrid, no oracle, no engine), so the table is not an output of the code it pins.

    s:    phi   = sum_k c_k (2 a_k / pi)^(3/4)            exp(-a_k r^2)
    p_j:  phi_j = sum_k c_k (128 a_k^5 / pi^3)^(1/4) d_j   exp(-a_k r^2),     d = r - R
    d/dx_i phi   = sum_k c_k N_k (-2 a_k d_i)              exp(-a_k r^2)
    d/dx_i phi_j = sum_k c_k N_k (delta_ij - 2 a_k d_i d_j) exp(-a_k r^2)

Shell data: the oxygen 2sp shell of STO-3G (Hehre, Stewart, Pople 1969), centre R chosen
off-axis.  Points: near the nucleus, on a nodal plane, mid-range, and far enough that the tightest primitive is
beyond any screening cutoff (a r^2 > 60) while the most diffuse one is not -- its true contribution there is
below 1e-27 of the shell's value, far under the comparison tolerance, so the table pins screened evaluators too.

    python tools/make_ao_pin_table.py > tests/golden/ao_pin_table.json
"""
import json

import mpmath as mp

mp.mp.dps = 50

EXPS = ["5.0331513", "1.1695961", "0.3803890"]
CS = ["-0.09996723", "0.39951283", "0.70011547"]
CP = ["0.15591627", "0.60768372", "0.39195739"]
CENTRE = ["0.1", "-0.2", "0.3"]
POINTS = [
    ["0.1", "-0.2", "0.3"], ["0.15", "-0.2", "0.3"], ["0.1", "0.4", "0.3"], ["-0.5", "0.3", "0.9"],
    ["1.2", "-1.1", "0.3"], ["0.1", "-0.2", "2.3"], ["-1.7", "1.9", "-0.8"], ["2.9", "2.2", "1.7"],
    ["-3.4", "-2.6", "3.1"], ["4.6", "-0.2", "0.3"],
]


def main():
    a = [mp.mpf(x) for x in EXPS]
    cs = [mp.mpf(x) for x in CS]
    cp = [mp.mpf(x) for x in CP]
    R = [mp.mpf(x) for x in CENTRE]
    ns = [(2 * ak / mp.pi) ** (mp.mpf(3) / 4) for ak in a]
    npn = [(128 * ak ** 5 / mp.pi ** 3) ** (mp.mpf(1) / 4) for ak in a]
    rows = []
    for P in POINTS:
        r = [mp.mpf(x) for x in P]
        d = [r[i] - R[i] for i in range(3)]
        r2 = sum(x * x for x in d)
        ex = [mp.e ** (-ak * r2) for ak in a]
        s_val = sum(cs[k] * ns[k] * ex[k] for k in range(3))
        s_grad = [sum(cs[k] * ns[k] * (-2 * a[k] * d[i]) * ex[k] for k in range(3)) for i in range(3)]
        p_val = [sum(cp[k] * npn[k] * d[j] * ex[k] for k in range(3)) for j in range(3)]
        p_grad = [[sum(cp[k] * npn[k] * ((1 if i == j else 0) - 2 * a[k] * d[i] * d[j]) * ex[k] for k in range(3))
                   for j in range(3)] for i in range(3)]   # [i = derivative direction][j = component]
        f = lambda x: mp.nstr(x, 20)
        rows.append({"point": P, "s": f(s_val), "s_grad": [f(x) for x in s_grad], "p": [f(x) for x in p_val],
                     "p_grad": [[f(x) for x in row] for row in p_grad]})
    print(json.dumps({"generator": "tools/make_ao_pin_table.py (mpmath, 50 digits, formulas only)",
                      "exps": EXPS, "coef_s": CS, "coef_p": CP, "centre": CENTRE,
                      "norm_s": "(2a/pi)^(3/4)", "norm_p": "(128 a^5/pi^3)^(1/4)", "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
