"""Stimuli -- mirror of tc_gan/stimuli.py (computed once per run on the host)."""
import numpy as np


def sigm(x, l=.1):
    return 1. / (1 + np.exp(-x / l))


def band(x, b, l=.1):
    """Smoothed top-hat of width b centred at 0.  stimuli.py:6-7."""
    return sigm(x + (b / 2), l) * sigm((b / 2) - x, l)


def input(bv, x, l=.1, c=[20.], o=[0.]):
    """
    [len(c) * len(o) * len(bv), 2N]: contrast-major, then offset, then bandwidth;
    the same profile for the E and the I half.  stimuli.py:9-10.
    """
    rows = []
    for con in c:
        for off in o:
            for b in bv:
                prof = band(x - off, b, l)
                rows.append(con * np.concatenate([prof, prof]))
    return np.array(rows)
