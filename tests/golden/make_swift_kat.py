"""Extracts the reference's SWIFT known-answer vectors (src/UnitTest.cpp:155-188 inputs,
:221-261 the 40 golden prices of `pricing_test`, :278-477 the 200 golden partials of
`gradient_test`, :478 the column order) into tests/golden/swift_kat.json.
Needs /root/reference (build container).  Run: python tests/golden/make_swift_kat.py"""
import json
import os
import re

SRC = "/root/reference/src/UnitTest.cpp"
HERE = os.path.dirname(os.path.abspath(__file__))
text = open(SRC).read()


def numbers(block):
    return [float(x) for x in re.findall(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?", block)]


def braces(name_regex, after=0):
    m = re.search(name_regex, text[after:])
    start = after + m.end()
    depth, i = 1, start
    while depth:
        depth += {"{": 1, "}": -1}.get(text[i], 0)
        i += 1
    return text[start:i - 1], i


exp_block, pos = braces(r"std::vector<double> expiries = \{")
strike_block, pos = braces(r"std::vector<std::vector<double>> strikes = \{", pos)
param_block, pos = braces(r"std::vector<swift_parameters> params=\{", pos)
p_block, pos = braces(r"double p\[5\]=\{", pos)
price_block, pos2 = braces(r"void pricing_test\(\)\{\s*double diff=0\.;\s*std::vector<ffloat> prices=\{")
grad_block, pos3 = braces(r"void gradient_test\(\)\{\s*double diff=0\.;\s*std::vector<ffloat> grad=\{")
idx_block, _ = braces(r"std::vector<ffloat> index_map=\{", pos3)

p_block = re.sub(r"//.*", "", p_block)
expiries = numbers(exp_block)
strikes = [numbers(b) for b in re.findall(r"\{([^{}]*)\}", strike_block)]
params = [numbers(b) for b in re.findall(r"\{([^{}]*)\}", param_block)]
out = {
    "source": "reference src/UnitTest.cpp:155-188,221-261,278-478 (golden values by E. Romo Grau's SWIFT)",
    "S": 1.0, "risk_free": 0.02,
    "hparams": numbers(p_block),          # v0, v_bar, rho, kappa, sigma
    "expiries": expiries, "strikes": strikes,
    "swift_parameters": params,           # m, exp2_m, sqrt_exp2_m, lower, upper, k_1, k_2, J
    "prices": numbers(price_block),
    "grad": numbers(grad_block),
    "grad_index_map": [int(x) for x in numbers(idx_block)],
}
assert len(expiries) == 8 and len(strikes) == 8 and all(len(s) == 5 for s in strikes)
assert len(params) == 8 and all(len(q) == 8 for q in params)
assert len(out["prices"]) == 40 and len(out["grad"]) == 200 and len(out["hparams"]) == 5
with open(os.path.join(HERE, "swift_kat.json"), "w") as f:
    json.dump(out, f, indent=1)
print("ok", out["hparams"], out["prices"][:2], out["grad"][:2], out["grad_index_map"])
