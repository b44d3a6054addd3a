import os

import numpy as np
import pytest

F = np.float32


def ref_test_vector(i, dim):
    """reference src/memory/retrieval.rs:2430-2438 test_vector (iterator sum is sequential in f32)."""
    v = np.ones(dim, F)
    v[i % dim] = F(v[i % dim] + F(F(0.5) * F(np.sqrt(F(dim)))))
    s = F(0)
    for x in v:
        s = F(s + F(x * x))
    norm = F(np.sqrt(s))
    return (v / norm).astype(F)


# GPU tests that were written after round 1's GPU budget was spent: compositions / re-parameterisations of code that IS
# verified, but they themselves have not run on a device yet.  They stay out of the default `-m gpu` run until they
# have (scripts/next_round_first_call.sh runs them with SHODH_UNVERIFIED=1); after a green run, drop the mark.
not_yet_run_on_gpu = pytest.mark.skipif(os.environ.get("SHODH_UNVERIFIED") != "1",
                                        reason="written after the round-1 GPU budget was spent: set SHODH_UNVERIFIED=1")
