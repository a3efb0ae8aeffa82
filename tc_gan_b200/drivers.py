"""
Learning-loop drivers -- mirror of tc_gan/drivers.py: `GANDriver` (:68-180), `SSNRejectionLimiter` (:214-255),
`WGANDiscLossLimiter` (:265-297), `BPTTWGANDriver` (:300-338), `BPTTcWGANDriver` (:341-352).  A driver records
learning statistics and parameters after every update, checks that the critic stays finite, and aborts with
the reference's `exit.json` reasons.
"""
import collections
import contextlib
import os
from logging import getLogger

import numpy as np

from . import execution, ssnode
from .recorders import (ConditionalTuningCurveStatsRecorder, DiscLearningRecorder, DiscParamStatsRecorder,
                        FlexGenParamRecorder, GenParamRecorder, LearningRecorder, LegacyLearningRecorder,
                        UpdateResult)

logger = getLogger(__name__)


def is_at_interval(step, interval):
    return interval > 0 and step % interval == 0


def net_isfinite(module):
    return all(bool(np.isfinite(p.detach().cpu().numpy()).all()) for p in module.parameters())


@contextlib.contextmanager
def recording_exit_reason(datastore):
    try:
        yield
    except KeyboardInterrupt:
        datastore.save_exit_reason(reason='keyboard_interrupt', good=False)
        raise
    except execution.KnownError:
        raise
    except Exception as err:
        datastore.save_exit_reason(reason='uncaught_exception', good=False, exception=str(err))
        raise
    else:
        datastore.save_exit_reason(reason='end_of_iteration', good=True)


def maybe_quit(datastore, JDS_fake, JDS_true, quit_JDS_threshold):
    JDS_fake = np.concatenate(JDS_fake).flatten()
    JDS_true = np.concatenate(JDS_true).flatten()
    JDS_distance = float(np.linalg.norm(JDS_fake - JDS_true))
    if quit_JDS_threshold > 0 and JDS_distance >= quit_JDS_threshold:
        datastore.dump_json(dict(reason='JDS_distance', JDS_distance=JDS_distance, good=False), 'exit.json')
        raise execution.KnownError('Exit simulation since (J, D, S)-distance (= {}) to the true parameter exceed '
                                   'threshold (= {}).'.format(JDS_distance, quit_JDS_threshold), exit_code=4)


def check_disc_param(datastore, discriminator, nnorms):
    isfinite_nnorms = np.isfinite(nnorms)
    if not isfinite_nnorms.all() and not net_isfinite(discriminator):
        datastore.dump_json(dict(reason='disc_param_has_nan', isfinite_nnorms=isfinite_nnorms.tolist(), good=False),
                            'exit.json')
        raise execution.KnownError("Discriminator parameter is not finite.", exit_code=3)


class SSNRejectionLimiter(object):
    """Watch the SSN rejection rate and terminate the driver if it is too large (tc_gan/drivers.py:214-255)."""

    def __init__(self, datastore, n_samples, rejection_limit=0.6, max_consecutive_exceedings=5):
        self.datastore, self.n_samples = datastore, n_samples
        self.rejection_limit, self.max_consecutive_exceedings = rejection_limit, max_consecutive_exceedings
        self._exceedings = 0

    def should_abort(self, rejections):
        if rejections / (rejections + self.n_samples) > self.rejection_limit:
            self._exceedings += 1
        else:
            self._exceedings = 0
        return self._exceedings > self.max_consecutive_exceedings

    def __call__(self, rejections):
        if self.should_abort(rejections):
            self.datastore.dump_json(dict(reason='too_many_rejections', good=False), 'exit.json')
            raise execution.KnownError("Too many rejections in fixed-point finder.", exit_code=4)

    @classmethod
    def from_driver(cls, driver):
        return cls(driver.datastore, n_samples=driver.gan.NZ)


class WGANDiscLossLimiter(object):

    def __init__(self, datastore, prob_limit=0.6, wild_disc_loss=10000, hist_length=50):
        self.datastore, self.prob_limit = datastore, prob_limit
        self.wild_disc_loss, self.hist_length = wild_disc_loss, hist_length
        self.dloss_hist = collections.deque(maxlen=hist_length)

    def prob_exceed(self):
        return np.mean(abs(np.asarray(self.dloss_hist) > self.wild_disc_loss))

    def should_abort(self, dloss):
        self.dloss_hist.append(dloss)
        return len(self.dloss_hist) == self.hist_length and self.prob_exceed() > self.prob_limit

    def __call__(self, dloss):
        if self.should_abort(dloss):
            self.datastore.dump_json(dict(reason='wild_disc_loss', good=False), 'exit.json')
            raise execution.KnownError("Too many wild discriminator losses.", exit_code=4)

    @classmethod
    def from_driver(cls, driver):
        return cls(driver.datastore)


def disc_loss_limiter(driver):
    if driver.gan.loss_type == 'WD':
        return WGANDiscLossLimiter.from_driver(driver)
    return lambda *_, **__: None


def dump_disc_param(discriminator, path):
    """Critic parameters as an ``.npz`` of arrays in parameter order (lasagne_toppings/param_file.py:30-60)."""
    np.savez(path, *[p.detach().cpu().numpy() for p in discriminator.parameters()])


class GANDriver(object):
    """Algorithm-independent bookkeeping of a GAN run: `iterate`, `post_disc_update`, `post_update`."""

    def make_learning_recorder(self):
        return LegacyLearningRecorder.from_driver(self)

    def make_generator_recorder(self):
        return GenParamRecorder.from_driver(self)

    def make_discparamstats_recorder(self):
        return DiscParamStatsRecorder.from_driver(self)

    def make_disclearning_recorder(self):
        return DiscLearningRecorder.from_driver(self)

    def __init__(self, gan, datastore, iterations=100, quiet=True, disc_param_save_interval=-1,
                 disc_param_template='last.npz', disc_param_save_on_error=False, quit_JDS_threshold=-1, **kwargs):
        self.gan, self.datastore = gan, datastore
        self.iterations, self.quiet = iterations, quiet
        self.disc_param_save_interval, self.disc_param_template = disc_param_save_interval, disc_param_template
        self.disc_param_save_on_error, self.quit_JDS_threshold = disc_param_save_on_error, quit_JDS_threshold
        self.__dict__.update(kwargs)

    def pre_loop(self):
        self.learning_recorder = self.make_learning_recorder()
        self.generator_recorder = self.make_generator_recorder()
        self.discparamstats_recorder = self.make_discparamstats_recorder()
        self.disclearning_recorder = self.make_disclearning_recorder()
        self.rejection_limiter = SSNRejectionLimiter.from_driver(self)
        self.disc_loss_limiter = disc_loss_limiter(self)

    def post_disc_update(self, gen_step, disc_step, Dloss, Daccuracy, SSsolve_time, gradient_time, model_info):
        self.disclearning_recorder.record(gen_step, disc_step, Dloss, Daccuracy, SSsolve_time, gradient_time,
                                          model_info.rejections, model_info.unused)
        nnorms = self.discparamstats_recorder.record(gen_step, disc_step)
        check_disc_param(self.datastore, self.gan.discriminator, nnorms)
        self.rejection_limiter(model_info.rejections)
        self.disc_loss_limiter(Dloss)

    def post_update(self, gen_step, update_result):
        self.learning_recorder.record(gen_step, update_result)
        jj, dd, ss = self.generator_recorder.record(gen_step)
        if is_at_interval(gen_step, self.disc_param_save_interval):
            dump_disc_param(self.gan.discriminator,
                            self.datastore.path('disc_param', self.disc_param_template.format(gen_step)))
        self.datastore.flush_all()
        maybe_quit(self.datastore, JDS_fake=list(map(np.exp, [jj, dd, ss])),
                   JDS_true=list(map(ssnode.DEFAULT_PARAMS.get, 'JDS')), quit_JDS_threshold=self.quit_JDS_threshold)

    def iterate(self, update_func):
        """Call ``update_func(gen_step) -> UpdateResult`` `iterations` times, recording after each."""
        self.pre_loop()
        with recording_exit_reason(self.datastore):
            for gen_step in range(self.iterations):
                try:
                    result = update_func(gen_step)
                except Exception:
                    if self.disc_param_save_on_error:
                        dump_disc_param(self.gan.discriminator, self.datastore.path('disc_param', 'post_error.npz'))
                    raise
                self.post_update(gen_step, result)


class BPTTWGANDriver(GANDriver):

    def make_learning_recorder(self):
        return LearningRecorder.from_driver(self)

    def make_generator_recorder(self):
        return FlexGenParamRecorder.from_driver(self)

    def run(self, gan):
        learning_it = gan.learning()
        state = {}

        def update_func(k):
            while True:
                info = next(learning_it)
                if info.is_discriminator:
                    self.post_disc_update(info.gen_step, info.disc_step, info.disc_loss, info.accuracy,
                                          info.gen_time, info.disc_time, ssnode.null_FixedPointsInfo)
                    state['disc_info'] = info
                else:
                    assert info.gen_step == k
                    disc_info = state['disc_info']
                    self.datastore.tables.saverow('TC_mean.csv', disc_info.xg.mean(axis=0).tolist()
                                                  + disc_info.xd.mean(axis=0).tolist())
                    return UpdateResult(info=info, disc_info=disc_info)

        self.iterate(update_func)


class BPTTcWGANDriver(BPTTWGANDriver):

    def __init__(self, *args, tc_stats_record_interval=100, **kwargs):
        super(BPTTcWGANDriver, self).__init__(*args, **kwargs)
        self.tc_stats_record_interval = tc_stats_record_interval

    def post_update(self, gen_step, update_result):
        if is_at_interval(gen_step, self.tc_stats_record_interval):
            self.tuning_curve_recorder.record(gen_step, update_result.disc_info)
        super(BPTTcWGANDriver, self).post_update(gen_step, update_result)

    def pre_loop(self):
        super(BPTTcWGANDriver, self).pre_loop()
        self.tuning_curve_recorder = ConditionalTuningCurveStatsRecorder.from_driver(self)
