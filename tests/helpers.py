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
