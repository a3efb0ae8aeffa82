"""
Probe / sub-sampling of rates into tuning curves -- mirror of
tc_gan/gradient_expressions/utils.py:27-149 (numpy and torch arrays).
"""
import numpy as np


def sample_sites_from_stim_space_impl(stim_locs, N, type=int):
    return ((stim_locs + 1) * (N - 1) / 2).astype(type)


def sample_sites_from_stim_space(stim_locs, N):
    """
    Neural indices of locations given in stimulus space [-1, 1].

    >>> sample_sites_from_stim_space([0, 0.5, 1], 101)
    [50, 75, 100]
    """
    stim_locs = np.asarray(stim_locs, dtype=float)
    if stim_locs.size and (stim_locs.min() < -1 or stim_locs.max() > 1):
        raise AssertionError('stimulus-space locations must lie in [-1, 1]')
    sample_sites = sample_sites_from_stim_space_impl(stim_locs, N)
    if len(sample_sites) != len(set(sample_sites)):
        raise ValueError(
            'Non-unique sample sites are specified.\n'
            'N (= {}) is not large enough for stim_locs (= {}) to'
            ' generate unique sample sites.'
            ' They generates sample_sites = {}'
            .format(N, list(stim_locs), list(sample_sites)))
    return [int(s) for s in sample_sites]


def subsample_neurons(rate_vector, sample_sites, track_offset_identity=False,
                      include_inhibitory_neurons=False, N=None, NZ=None, NB=None):
    """
    Rates (NZ, NB, 2N) -> what the critic sees.  Probed neurons are `sample_sites` (excitatory
    indices; with include_inhibitory_neurons also their inhibitory partners at +N).  The probes of
    one network either stay together, (NZ, NB * n_probes) when track_offset_identity, or become
    separate samples, (NZ * n_probes, NB).  numpy arrays and torch tensors.

    >>> r = np.tile(np.arange(14), (5, 2, 1))
    >>> subsample_neurons(r, [2, 3, 4]).shape
    (15, 2)
    >>> subsample_neurons(r, [2, 3, 4], True)[0].tolist()
    [2, 3, 4, 2, 3, 4]
    """
    nz, nb, width = rate_vector.shape
    n_sites = width // 2 if N is None else N
    if (NZ is not None and NZ != nz) or (NB is not None and NB != nb) or width != 2 * n_sites:
        raise AssertionError('rate_vector must have shape (NZ, NB, 2N)')
    probes = [int(p) for p in sample_sites]
    if not probes or min(probes) < 0 or max(probes) >= n_sites:
        raise AssertionError('sample_sites must lie in [0, N)')
    if include_inhibitory_neurons:
        probes += [p + n_sites for p in probes]
    picked = rate_vector[:, :, probes]                              # (NZ, NB, n_probes)
    if track_offset_identity:
        return picked.reshape((nz, nb * len(probes)))
    moved = picked.swapaxes(1, 2) if isinstance(picked, np.ndarray) else picked.transpose(1, 2)
    return moved.reshape((nz * len(probes), nb))
