"""
"True data" tuning curves through the fixed-point solver -- mirror of tc_gan/networks/dataset.py:28-71.
"""
from logging import getLogger

import numpy as np

from .. import ssnode

logger = getLogger(__name__)

dataset_provider_choices = ('ssnode',)


def log_descriptive_stats(name, data):
    logger.info('Summary statistics of %s:', name)
    for label, value in (('mean', data.mean()), ('std ', data.std()), ('min ', data.min()),
                         ('25% ', np.percentile(data, 25)), ('50% ', np.percentile(data, 50)),
                         ('75% ', np.percentile(data, 75)), ('max ', data.max())):
        logger.info('  %s: %s', label, value)


def dataset_by_ssnode(num_sites, bandwidths, contrasts, truth_size, truth_seed, sample_sites,
                      include_inhibitory_neurons, true_ssn_options={}):
    """
    `truth_size` networks of the true SSN sampled with `ssnode.sample_tuning_curves`
    (asym_power transfer function, networks that reach rate_stop_at = 200 rejected and re-drawn,
    dt = 5e-4, max_iter = 100000 as the reference), probed at `sample_sites`.
    Returns an array of shape ``(truth_size, n_contrasts * n_bandwidths * n_cell_types * n_probes)``.
    """
    options = dict(dt=5e-4, max_iter=100000, io_type='asym_power', rate_stop_at=200)
    options.update(true_ssn_options)
    data, (_, _, fpinfo) = ssnode.sample_tuning_curves(
        sample_sites=sample_sites, NZ=truth_size, seed=truth_seed, bandwidths=bandwidths,
        contrast=contrasts, N=num_sites, track_offset_identity=True,
        include_inhibitory_neurons=include_inhibitory_neurons, **options)
    data = np.array(data.T)
    log_descriptive_stats('tuning curve dataset', data)
    logger.info('  rejections: %s (rate %s), error codes: %r', fpinfo.rejections,
                fpinfo.rejections / (fpinfo.rejections + len(data)), fpinfo.counter)
    return data
