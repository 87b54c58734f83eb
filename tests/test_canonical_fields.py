"""The neutral workload generator (tools/canonical.py, used by bench.py's product arm) against the oracle's own
generator, bit for bit -- and through it against the FNV anchors of SURVEY.md Appendix D (tests/test_oracle_pin.py)."""
import numpy as np
import pytest

from tools import canonical


@pytest.mark.parametrize("n", [64, 256, 1000])
def test_fields_equal_the_oracle_generator(sfo, n):
    got, want = canonical.fields(n), sfo.canonical_fields(n)
    for name, a, b in zip(("d", "u", "v", "sd", "su", "sv"), got, want):
        assert np.array_equal(a.view(np.int32), b.view(np.int32)), name


def test_row_slices_are_slices_of_the_full_fields():
    n = 300
    full = canonical.fields(n, threads=1)
    for r0, r1 in ((0, 17), (100, 260), (299, 300)):
        part = canonical.rows(n, r0, r1)
        for a, b in zip(part, full):
            assert a.shape == (r1 - r0, n) and np.array_equal(a, b[r0:r1])
