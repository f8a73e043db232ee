import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def oracle_kinds():
    """Oracles available in this environment: the compiled reference where oracle/_ref exists,
    and always the C port."""
    from oracle import oracle
    kinds = ["port"]
    if oracle.have_ref("trws"):
        kinds.insert(0, "reference")
    return kinds


def best_oracle():
    return oracle_kinds()[0]


def trws_oracle(pr, maxiter, relgap=0.0, kind=None):
    from oracle import oracle
    return oracle.trws_solve(pr["kernel"], pr["unary"].T, (pr["connectivity"] - 1).T, pr["q"].T, pr["qprim"].T,
                             pr["alphas"], pr["tol"], maxiter, relgap, kind=kind or best_oracle())
