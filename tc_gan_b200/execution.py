"""
Run directory ("datastore") of a learning run -- mirror of tc_gan/execution.py:110-365: CSV tables
(`DataTables`), HDF5 tables in ``store.hdf5`` / dedicated ``<table>.hdf5`` files (`HDF5Tables`, when h5py is
importable), ``info.json`` (`pre_learn`) and ``exit.json`` (`save_exit_reason`), so that the reference's
loaders (tc_gan/loaders/datastore_loader.py:58-75: ``<table>.csv`` first, then the HDF5 files) read a run of
this package.
"""
import json
import os
import sys
import time

import numpy as np


class KnownError(Exception):
    """An exception that is expected to happen; carries the process exit code (tc_gan/execution.py:24-35)."""

    def __init__(self, message, exit_code=1):
        self.exit_code = exit_code
        super(KnownError, self).__init__(message)


def have_h5py():
    try:
        import h5py  # noqa: F401
        return True
    except ImportError:
        return False


class DataTables(object):
    """Append-only CSV files, one per table name (tc_gan/execution.py:110-157)."""

    def __init__(self, directory):
        self.directory = directory
        self._files = {}

    def _open(self, name):
        return open(os.path.join(self.directory, name), 'w')

    def _get_file(self, name):
        if name not in self._files:
            self._files[name] = self._open(name)
        return self._files[name]

    def saverow(self, name, row, echo=False, flush=False):
        if isinstance(row, (list, tuple)):
            row = ','.join(map(str, row))
        file = self._get_file(name)
        file.write(row)
        file.write('\n')
        if flush:
            file.flush()
        if echo:
            print(row)

    def flush_all(self):
        for file in self._files.values():
            file.flush()

    def __enter__(self):
        return self

    def __exit__(self, type, value, traceback):
        for name, file in self._files.items():
            try:
                file.close()
            except Exception as err:
                print('Error while closing', name, err, 'ignoring...')


class HDF5Tables(object):
    """Extensible 1-D compound datasets, shared ``store.hdf5`` or ``<name>.hdf5`` (tc_gan/execution.py:160-198)."""

    shared_filename = 'store.hdf5'

    def __init__(self, h5):
        self.h5 = h5
        self._datasets = {}

    def _create_dataset(self, name, dtype, dedicated=False):
        file = self.h5.open(name + '.hdf5' if dedicated else self.shared_filename)
        return file.create_dataset(name, (0,), maxshape=(None,), dtype=dtype), file

    def _get_dataset(self, name, *args, **kwds):
        if name not in self._datasets:
            self._datasets[name] = self._create_dataset(name, *args, **kwds)
        return self._datasets[name]

    def create_table(self, name, dtype, dedicated=False):
        assert name not in self._datasets
        self._get_dataset(name, dtype, dedicated)

    def saverow(self, name, row, echo=False, flush=False):
        dataset, file = self._get_dataset(name, row.dtype)
        dataset.resize((len(dataset) + 1,))
        dataset[-1] = row
        if flush:
            file.flush()
        if echo:
            print(*row.tolist(), sep=',')


class HDF5Store(object):

    def __init__(self, datastore):
        self.datastore = datastore
        self.tables = HDF5Tables(self)
        self._files = {}

    def _open(self, filename):
        import h5py
        return self.datastore.enter_context(h5py.File(os.path.join(self.datastore.directory, filename), 'w'))

    def open(self, filename):
        if filename not in self._files:
            self._files[filename] = self._open(filename)
        return self._files[filename]

    def flush_all(self):
        for file in self._files.values():
            file.flush()


def makedirs_exist_ok(name):
    os.makedirs(name, exist_ok=True)


class DataStore(object):
    """tc_gan/execution.py:222-267.  `table_format`: 'hdf5' (the reference's default; needs h5py), 'csv', or
    'auto' = hdf5 when h5py is importable, else csv (both are read by the reference's loaders)."""

    def __init__(self, directory, table_format='auto'):
        self.directory = directory
        makedirs_exist_ok(directory)
        if table_format == 'auto':
            table_format = 'hdf5' if have_h5py() else 'csv'
        if table_format not in ('hdf5', 'csv'):
            raise ValueError('Unknown table_format: {}'.format(table_format))
        self.table_format = table_format
        self.tables = DataTables(directory)
        self.h5 = HDF5Store(self)
        self.exit_hooks = []

    def path(self, *subpaths):
        newpath = os.path.join(self.directory, *subpaths)
        makedirs_exist_ok(os.path.dirname(newpath))
        return newpath

    def dump_json(self, obj, filename):
        with open(self.path(filename), 'w') as fp:
            json.dump(obj, fp)

    def save_exit_reason(self, reason, good, **kwargs):
        self.dump_json(dict(reason=reason, good=good, **kwargs), 'exit.json')

    def flush_all(self):
        self.tables.flush_all()
        self.h5.flush_all()

    def __repr__(self):
        return '<DataStore: {}>'.format(self.directory)

    def enter_context(self, context_manager):
        ret = context_manager.__enter__()
        self.exit_hooks.append(context_manager.__exit__)
        return ret

    def __enter__(self):
        self.enter_context(self.tables)
        return self

    def __exit__(self, *exc):
        run_exit_hooks(self.exit_hooks, exc)


def run_exit_hooks(exit_hooks, exc=(None, None, None)):
    if not exit_hooks:
        return
    try:
        exit_hooks[0](*exc)
    finally:
        run_exit_hooks(exit_hooks[1:], exc)


def format_datastore(datastore_template, run_config):
    """
    >>> format_datastore('alpha={alpha}_L={layers_str}', dict(alpha=10, layers=[128, 64]))
    'alpha=10_L=128_64'
    """
    return datastore_template.format(layers_str='_'.join(map(str, run_config.get('layers', []))), **run_config)


def get_meta_info():
    import platform
    info = dict(python=sys.version, platform=platform.platform(), argv=sys.argv, time=time.strftime('%Y-%m-%d %H:%M:%S'),
                numpy=np.__version__)
    try:
        import torch
        info['torch'] = torch.__version__
    except ImportError:
        pass
    return info


def _jsonable(obj):
    if isinstance(obj, dict):
        return {str(k): _jsonable(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_jsonable(v) for v in obj]
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, (np.integer,)):
        return int(obj)
    if isinstance(obj, (np.floating,)):
        return float(obj)
    return obj


def pre_learn(datastore=None, datastore_template='logfiles/{IO_type}', extra_info={}, preprocess=None, **run_config):
    """Create the run directory and write ``info.json`` = {run_config, extra_info, meta_info}
    (tc_gan/execution.py:318-346)."""
    if preprocess:
        preprocess(run_config)
    if not datastore:
        datastore = format_datastore(datastore_template, run_config)
    makedirs_exist_ok(datastore)
    with open(os.path.join(datastore, 'info.json'), 'w') as fp:
        json.dump(dict(run_config=_jsonable(run_config), extra_info=_jsonable(extra_info), meta_info=get_meta_info()), fp)
    run_config['datastore'] = datastore
    return run_config


def do_learning(learn, run_config, extra_info={}, preprocess=None, table_format='auto'):
    """``learn(datastore=DataStore, **run_config)`` after `pre_learn` (tc_gan/execution.py:349-365)."""
    run_config = pre_learn(extra_info=extra_info, preprocess=preprocess, **run_config)
    with DataStore(run_config.pop('datastore'), table_format=table_format) as datastore:
        return learn(datastore=datastore, **run_config)
