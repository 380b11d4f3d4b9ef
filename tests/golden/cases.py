"""Golden cases shared by the generator (make_golden.py) and the tests."""

DEFAULT = (0.04, 0.04, -0.7, 2.0, 0.5)
STIFF = (0.04, 0.04, -0.95, 20.0, 1.0)

# name -> (payoff, expiries, strikes per chain, steps, params, n_sims)
PRICE_CASES = {
    "cfg1_asian_252": (0, [1.0], [[100.0]], 252, DEFAULT, 4000),
    "cfg2_european_252": (1, [1.0], [[100.0]], 252, DEFAULT, 4000),
    "cfg4_asian_1024_quirk": (0, [1.0], [[100.0]], 1024, DEFAULT, 1000),
    "asian_two_chains": (0, [0.5, 1.0], [[90.0, 100.0], [100.0, 110.0]], 252, DEFAULT, 3000),
    "european_same_step_chains": (1, [0.25, 0.26, 1.0], [[90.0, 100.0], [100.0, 110.0], [95.0]],
                                  100, DEFAULT, 3000),
    "asian_same_step_chains": (0, [0.25, 0.2501, 0.2502, 0.6], [[95.0], [100.0, 101.0], [99.0], [100.0]],
                               50, DEFAULT, 3000),
    "cfg3_asian_chain_8x8": (0, [0.25 * k for k in range(1, 9)],
                             [[70.0 + 60.0 * j / 7 for j in range(8)]] * 8, 252, DEFAULT, 1500),
    "cfg5_stiff_asian": (0, [10.0], [[70.0, 100.0, 130.0]], 2520, STIFF, 300),
    "low_vol_of_vol_exp_branch": (0, [1.0], [[100.0]], 64, (0.01, 0.02, -0.3, 0.5, 1.5), 3000),
    "asian_365": (0, [1.0], [[100.0]], 365, DEFAULT, 2000),
}

# (size, seed, pattern) for the RNG wrapper: pattern 'u'/'g' repeated `reps` times
RNG_CASES = {
    "seed1_alternate": (256, 1, "ug", 700),      # crosses several refills of both buffers
    "seed2_g_only": (128, 2, "g", 600),
    "seed4_mostly_g": (128, 4, "ggggggu", 200),
    "seed128_u_only": (256, 128, "u", 600),
}
