"""Mirror of tc_gan/networks/utils.py:1-70."""

gridified_tc_axes = ('sample', 'cell_type', 'norm_probe', 'contrast', 'bandwidth')
"""Names of the axes of the array returned by `gridify_tc_samples`."""

sampled_tc_axes = ('sample', 'contrast', 'bandwidth', 'cell_type', 'norm_probe')
"""Names of the axes of the array accepted by `gridify_tc_samples`."""


def gridify_tc_samples(data, num_contrasts, num_bandwidths, num_cell_types, num_probes):
    """
    Tuning-curve `data` as `subsample_neurons(..., track_offset_identity=True)` lays it out --
    per sample: contrast-major, then bandwidth, then cell type, then probe -- to the grid
    ``(sample, cell_type, norm_probe, contrast, bandwidth)`` (tc_gan/networks/utils.py:10-70).

    >>> import numpy as np
    >>> shape = (11, 5, 7, 2, 3)
    >>> data = np.arange(np.prod(shape)).reshape(shape)
    >>> grid = gridify_tc_samples(data.reshape(11, -1), num_contrasts=5, num_bandwidths=7,
    ...                           num_cell_types=2, num_probes=3)
    >>> grid.shape
    (11, 2, 3, 5, 7)
    >>> bool(grid[4, 1, 2, 3, 6] == data[4, 3, 6, 1, 2])
    True
    """
    grid = data.reshape((len(data), num_contrasts, num_bandwidths, num_cell_types, num_probes))
    return grid.transpose((0, 3, 4, 1, 2))
