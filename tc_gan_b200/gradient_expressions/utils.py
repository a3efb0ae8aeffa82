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
    stim_locs = np.asarray(stim_locs)
    assert all(stim_locs >= -1)
    assert all(stim_locs <= 1)
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
    (NZ, NB, 2N) rates -> (NZ * n_sites, NB), or (NZ, NB * n_sites) when
    track_offset_identity.  Works on numpy arrays and torch tensors.

    >>> r = np.tile(np.arange(14), (5, 2, 1))
    >>> subsample_neurons(r, [2, 3, 4]).shape
    (15, 2)
    >>> subsample_neurons(r, [2, 3, 4], True)[0].tolist()
    [2, 3, 4, 2, 3, 4]
    """
    NZ_, NB_, TN_ = rate_vector.shape
    NZ = NZ_ if NZ is None else NZ
    NB = NB_ if NB is None else NB
    N = TN_ // 2 if N is None else N
    assert (NZ_, NB_, TN_) == (NZ, NB, 2 * N)
    assert 0 <= min(sample_sites)
    assert max(sample_sites) < N
    sample_sites = list(sample_sites)
    if include_inhibitory_neurons:
        sample_sites = sample_sites + [s + N for s in sample_sites]
    subsample = rate_vector[:, :, sample_sites]
    if track_offset_identity:
        return subsample.reshape((NZ, -1))
    if isinstance(subsample, np.ndarray):
        return subsample.swapaxes(1, 2).reshape((-1, NB))
    return subsample.transpose(1, 2).reshape((-1, NB))
