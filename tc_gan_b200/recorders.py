"""
Learning-statistics tables -- mirror of tc_gan/recorders.py:58-397 (same table names, column names and
dtypes).  A recorder writes through the datastore in its `table_format`: HDF5 compound datasets exactly as the
reference (h5py), or ``<tablename>.csv`` with a header line, which the reference's `DataStoreLoader.default_load`
reads first (tc_gan/loaders/datastore_loader.py:58-75).
"""
import collections
import itertools

import numpy as np


class UpdateResult(object):
    """Result of a generator update (attributes as tc_gan/recorders.py:11-56)."""

    dynamics_penalty = np.nan

    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)


class BaseRecorder(object):

    dedicated = False

    def __init__(self, datastore, quiet=True):
        self.datastore = datastore
        self.quiet = quiet

    @property
    def column_names(self):
        return self.dtype.names

    def record(self, *row):
        self._saverow(row)

    def _use_hdf5(self):
        return getattr(self.datastore, 'table_format', 'csv') == 'hdf5'

    def _saverow(self, row):
        row = list(row)
        assert len(row) == len(self.column_names), (len(row), self.column_names)
        if self._use_hdf5():
            typed_row = np.array(tuple(row), dtype=self.dtype)
            self.datastore.h5.tables.saverow(self.tablename, typed_row, echo=not self.quiet)
        else:
            typed = np.array(tuple(row), dtype=self.dtype).tolist()        # cast as the HDF5 table would
            self.datastore.tables.saverow(self.tablename + '.csv', list(typed), echo=not self.quiet)

    def write_header(self):
        if self._use_hdf5():
            self.datastore.h5.tables.create_table(self.tablename, self.dtype, dedicated=self.dedicated)
        else:
            self.datastore.tables.saverow(self.tablename + '.csv', list(self.column_names))

    @classmethod
    def make(cls, *args, **kwargs):
        self = cls(*args, **kwargs)
        self.write_header()
        return self

    @classmethod
    def from_driver(cls, driver):
        return cls.make(driver.datastore)


HDF5Recorder = CSVRecorder = BaseRecorder


class LearningRecorder(BaseRecorder):

    tablename = 'learning'
    dtype = np.dtype([
        ('gen_step', 'uint32'), ('Gloss', 'double'), ('Dloss', 'double'), ('Daccuracy', 'double'),
        ('gen_forward_time', 'double'), ('gen_train_time', 'double'), ('disc_time', 'double'),
        ('rate_penalty', 'double'), ('dynamics_penalty', 'double'),
    ])

    def record(self, gen_step, update_result):
        info, disc_info = update_result.info, update_result.disc_info
        self._saverow([gen_step, info.gen_loss, disc_info.disc_loss, disc_info.accuracy, info.gen_forward_time,
                       info.gen_train_time, info.disc_time, disc_info.rate_penalty, disc_info.dynamics_penalty])

    @classmethod
    def from_driver(cls, driver):
        return cls.make(driver.datastore, quiet=driver.quiet)


class LegacyLearningRecorder(BaseRecorder):
    """`learning` table of the fixed-point GAN (tc_gan/recorders.py:364-397)."""

    tablename = 'learning'
    dtype = np.dtype([
        ('gen_step', 'uint32'), ('Gloss', 'double'), ('Dloss', 'double'), ('Daccuracy', 'double'),
        ('SSsolve_time', 'double'), ('gradient_time', 'double'), ('model_convergence', 'uint32'),
        ('model_unused', 'uint32'), ('rate_penalty', 'double'), ('dynamics_penalty', 'double'),
    ])

    def record(self, gen_step, update_result):
        u = update_result
        self._saverow([gen_step, u.Gloss, u.Dloss, u.Daccuracy, u.SSsolve_time, u.gradient_time,
                       u.model_info.rejections, u.model_info.unused, u.rate_penalty, u.dynamics_penalty])

    @classmethod
    def from_driver(cls, driver):
        return cls.make(driver.datastore, quiet=driver.quiet)


class DiscLearningRecorder(BaseRecorder):

    tablename = 'disc_learning'
    dtype = np.dtype([
        ('gen_step', 'uint32'), ('disc_step', 'uint32'), ('Dloss', 'double'), ('Daccuracy', 'double'),
        ('SSsolve_time', 'double'), ('gradient_time', 'double'), ('model_convergence', 'uint32'),
        ('model_unused', 'uint32'),
    ])


def _genparam_names():
    """
    >>> _genparam_names()[:5]
    ('J_EE', 'J_EI', 'J_IE', 'J_II', 'D_EE')
    """
    return tuple(p + s for p in 'JDS' for s in ('_EE', '_EI', '_IE', '_II'))


def gen_param_dtype(names):
    return [('gen_step', 'uint32')] + [(n, 'double') for n in names]


class GenParamRecorder(BaseRecorder):

    tablename = 'generator'
    dtype = np.dtype(gen_param_dtype(_genparam_names()))

    def __init__(self, datastore, gan):
        self.gan = gan
        super(GenParamRecorder, self).__init__(datastore)

    def record(self, gen_step):
        jj, dd, ss = self.gan.get_gen_param()
        self._saverow([gen_step] + list(np.concatenate([jj, dd, ss]).flat))
        return [jj, dd, ss]

    @classmethod
    def from_driver(cls, driver):
        return cls.make(driver.datastore, driver.gan)


class FlexGenParamRecorder(GenParamRecorder):
    """`generator` table with the heteroin parameter V when present (tc_gan/recorders.py:262-276)."""

    def __init__(self, *args, **kwargs):
        super(FlexGenParamRecorder, self).__init__(*args, **kwargs)
        self.dtype = np.dtype(gen_param_dtype(self.gan.gen.get_flat_param_names()))

    def record(self, gen_step):
        self._saverow([gen_step] + list(self.gan.gen.get_flat_param_values()))
        return self.gan.get_gen_param()


class DiscParamStatsRecorder(BaseRecorder):
    """Normalised norms ||p|| / p.size of every critic parameter (tc_gan/recorders.py:279-318); the columns are
    named ``<param>.nnorm.<k>`` with Lasagne's parameter names ('W', 'b', ...)."""

    tablename = 'disc_param_stats'

    def __init__(self, datastore, discriminator):
        self.discriminator = discriminator
        super(DiscParamStatsRecorder, self).__init__(datastore)
        names = [{'weight': 'W', 'bias': 'b'}.get(n.rsplit('.', 1)[-1], n.rsplit('.', 1)[-1])
                 for n, _ in discriminator.named_parameters()]
        self.dtype = np.dtype([('gen_step', 'uint32'), ('disc_step', 'uint32')] +
                              [(name, 'double') for name in self.disc_param_unique_names(names)])

    @staticmethod
    def disc_param_unique_names(names):
        counter = collections.Counter()
        for n in names:
            yield '{}.nnorm.{}'.format(n, counter[n])
            counter[n] += 1

    def record(self, gen_step, disc_step):
        nnorms = [float(p.detach().norm()) / p.numel() for p in self.discriminator.parameters()]
        self._saverow([gen_step, disc_step] + nnorms)
        return nnorms

    @classmethod
    def from_driver(cls, driver):
        return cls.make(driver.datastore, driver.gan.discriminator)


class ConditionalTuningCurveStatsRecorder(BaseRecorder):
    """Per-condition mean / variance of true and fake tuning curves (tc_gan/recorders.py:321-361)."""

    tablename = 'tc_stats'
    dedicated = True

    def __init__(self, datastore, num_bandwidths):
        super(ConditionalTuningCurveStatsRecorder, self).__init__(datastore)
        self.num_bandwidths = num_bandwidths
        self.dtype = np.dtype([
            ('gen_step', 'uint32'), ('is_fake', 'b'), ('contrast', 'double'), ('norm_probe', 'double'),
            ('cell_type', 'uint16'), ('count', 'uint32'),
        ] + [('mean_{}'.format(i), 'double') for i in range(num_bandwidths)]
          + [('var_{}'.format(i), 'double') for i in range(num_bandwidths)])

    @staticmethod
    def analyze(tuning_curves, conditions):
        key = lambda i: tuple(conditions[i])
        indices = sorted(range(len(conditions)), key=key)
        for cond, group in itertools.groupby(indices, key=key):
            tc = tuning_curves[list(group)]
            yield list(cond) + [len(tc)] + list(tc.mean(axis=0)) + list(tc.var(axis=0))

    def record(self, gen_step, info):
        for is_fake, x, c in [(0, info.xd, info.cd), (1, info.xg, info.cg)]:
            for cond_stats in self.analyze(x, c):
                self._saverow([gen_step, is_fake] + list(cond_stats))

    @classmethod
    def from_driver(cls, driver):
        return cls.make(driver.datastore, len(driver.gan.bandwidths))
