"""CPU tests of the host-side mirror of the reference interface: flag system, Dataset index generation (vectorised
vs the reference's per-row loop restated in the oracle), IndexSet, LR schedules, initialisers, weight container."""
import os

import numpy as np
import pandas as pd
import pytest

import lfm_oracle as orc
from lfm_quant_b200.scripts import base_config, configs
from lfm_quant_b200.scripts.data_processing import Dataset, IndexSet
from lfm_quant_b200.scripts.model_utils.initializers import Initializer
from lfm_quant_b200.scripts.model_utils.optimizers import Optimizers
from lfm_quant_b200.scripts.synthetic import write_open_dataset, write_system_test_conf


@pytest.fixture()
def fresh_flags():
    configs.reset()
    yield
    configs.reset()


def test_flag_defaults_match_reference_schema(fresh_flags):
    c = base_config.get_configs([])
    # spot checks against scripts/lfm_quant.py:23-106
    assert c.nn_type == 'RNNPointEstimate' and c.optimizer == 'Adadelta' and c.learning_rate == 0.6
    assert c.num_layers == 2 and c.num_hidden == 64 and c.batch_size == 256 and c.stride == 12
    assert c.max_grad_norm == 50.0 and c.max_norm == 3 and c.seed == 521 and c.train is True
    assert c.target_lambda == 0.5 and c.rnn_lambda == 0.7 and c.lr_schedule == 'ExponentialDecay'
    assert c.forecast_steps_weights == [1.0]
    assert c.piecewise_lr_boundaries == [4000.0, 5500.0, 5500.0]
    assert c.piecewise_lr_values == [0.5, 0.1, 0.05, 0.1]
    assert len(base_config.SCHEMA) >= 85


def test_boolean_flag_forms_and_unknown_flags(fresh_flags):
    c = base_config.get_configs(['--train=False', '--nolog_squasher', '--UQ', '--aux_masking', 'true', '--bogus', '1'])
    assert c.train is False and c.log_squasher is False and c.UQ is True and c.aux_masking is True
    with pytest.raises(AttributeError):
        c.bogus


def test_config_file_and_override_order(fresh_flags, tmp_path):
    f = tmp_path / 'm.conf'
    f.write_text('--num_hidden 128\n--optimizer Adam --min_years 3\n--max_years 5 --stride 12\n')
    c = base_config.get_configs(['--config', str(f), '--num_hidden', '32'])
    assert c.num_hidden == 32 and c.optimizer == 'Adam'
    assert c.min_unrollings == 3 and c.max_unrollings == 5          # lfm_quant.py:116-123
    c.foo = 7                                                         # settable bag (configs.py:53-56)
    assert c.foo == 7


def test_lr_schedules(fresh_flags):
    c = base_config.get_configs(['--lr_decay', '0.5', '--decay_steps', '10'])
    sch = Optimizers(c).get_learning_rate()
    assert sch(9) == pytest.approx(0.6) and sch(10) == pytest.approx(0.3) and sch(25) == pytest.approx(0.15)
    c.lr_schedule = 'PolynomialDecay'
    sch = Optimizers(c).get_learning_rate()
    assert sch(5) == pytest.approx(orc.learning_rate(5, lr_schedule='PolynomialDecay', learning_rate=0.6, decay_steps=10,
                                                     end_learning_rate=0.01, decay_power=0.5))
    c.lr_schedule = 'PiecewiseConstantDecay'
    sch = Optimizers(c).get_learning_rate()
    assert sch(4000) == 0.5 and sch(4001) == 0.1 and sch(9999) == 0.1
    c.lr_schedule = 'Nope'
    with pytest.raises(ValueError):
        Optimizers(c)
    c.lr_schedule = 'ExponentialDecay'
    c.optimizer = 'Lion'
    with pytest.raises(ValueError):
        Optimizers(c).get_optimizer()


def test_index_set_shuffle_and_batch():
    n = 25
    idx = np.arange(n * 3).reshape(n, 3).astype(np.int32)
    s = IndexSet(idx, idx + 1000, np.arange(n * 3).reshape(n, 3).astype('S'))
    a = s.shuffle(10, seed=3)
    b = s.shuffle(10, seed=3)
    np.testing.assert_array_equal(a.inp_idx, b.inp_idx)
    assert sorted(a.inp_idx[:, 0].tolist()) == idx[:, 0].tolist() and not np.array_equal(a.inp_idx, idx)
    np.testing.assert_array_equal(a.tar_idx, a.inp_idx + 1000)        # rows stay zipped
    batches = list(a.batch(8))
    assert [x[0].shape[0] for x in batches] == [8, 8, 8, 1] and len(a.batch(8)) == 4


@pytest.fixture(scope='module')
def dataset_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp('data')
    write_open_dataset(str(d / 'open-dataset.dat'), n_keys=14, n_months=400, seed=5)
    return d


def _dataset(dataset_dir, tmp_path, extra=()):
    configs.reset()
    conf = tmp_path / 'system-test.conf'
    write_system_test_conf(str(conf), str(dataset_dir), str(tmp_path / 'exp'))
    c = base_config.get_configs(['--config', str(conf)] + list(extra))
    return c, Dataset(c)


@pytest.mark.parametrize('extra', [(), ('--train=False',), ('--min_unrollings', '3', '--max_unrollings', '6', '--stride', '6',
                                                           '--forecast_n', '6')])
def test_window_index_matches_reference_loop(dataset_dir, tmp_path, extra):
    c, ds = _dataset(dataset_dir, tmp_path, extra)
    ds._create_index()
    keys = ds._keys
    dates = pd.DatetimeIndex(ds._dates)
    inp, tar, rows = orc.create_window_index(
        keys, ds.table[:, ds._active_idx], dates, train=c.train, stride=c.stride, forecast_n=c.forecast_n,
        min_unrollings=c.min_unrollings, max_unrollings=c.max_unrollings, start_date=ds.start_date,
        end_date=ds.end_date, last_train_date=ds.end_date - pd.DateOffset(months=c.stride))
    assert inp.shape[0] > 50
    if c.train:
        tr = np.isin(keys[rows], np.asarray(ds._train_gvkeys, dtype=keys.dtype))
        np.testing.assert_array_equal(ds._dataset['train_X'], inp[tr])
        np.testing.assert_array_equal(ds._dataset['train_Y'], tar[tr])
        np.testing.assert_array_equal(ds._dataset['valid_X'], inp[~tr])
        np.testing.assert_array_equal(ds._dataset['valid_Y'], tar[~tr])
        assert ds._dataset['valid_X'].shape[0] > 0
        assert ds._meta['train'].shape == (int(tr.sum()), 3)
    else:
        np.testing.assert_array_equal(ds._dataset['test_X'], inp)
        np.testing.assert_array_equal(ds._dataset['test_Y'], tar)
        assert (tar[:, 1] == inp[:, 1]).any()          # windows whose target does not exist (:275-279)
    configs.reset()


def test_dataset_attributes_and_scaler(dataset_dir, tmp_path):
    c, ds = _dataset(dataset_dir, tmp_path)
    assert ds.n_inputs == 32 and ds.n_outputs == 16 and ds.seq_len == 20 and ds.target_index == 3
    assert ds._cols[:3] == ['date', 'gvkey', 'active'] and ds._cols[-1] == 'mrkcap'
    assert ds.scale_inp_col_ids == list(range(32)) and ds._seq_norm_idx == 35
    ds.generate_dataset()
    sp = ds.scaling_params
    assert set(sp) == {'center', 'scale'} and sp['center'].shape == (32,) and (sp['scale'] > 0).all()
    assert os.path.isfile(os.path.join(c.experiments_dir, c.model_dir, 'scales.dat'))
    # second Dataset re-loads the pickle instead of re-fitting
    ds2 = Dataset(c)
    ds2.generate_dataset()
    np.testing.assert_array_equal(ds2.scaling_params['center'], sp['center'])
    x = np.array([-3.0, 0.0, 2.5])
    np.testing.assert_allclose(ds.reverse_log_squasher(ds.log_squasher(x)), x)
    configs.reset()


def test_initializer_families(fresh_flags):
    c = base_config.get_configs(['--init_scale', '0.25'])
    specs = list(zip(orc.param_names(1), orc.param_shapes(1, 8, 3, 12)))
    w = Initializer(c).initial_weights(specs)
    assert np.abs(w[0]).max() <= 0.25 and w[0].dtype == np.float32
    np.testing.assert_allclose(w[1] @ w[1].T, np.eye(12), atol=1e-5)            # orthogonal recurrent kernel
    assert (w[2][12:24] == 1).all() and w[2].sum() == 12                        # unit forget bias
    assert (w[3] == 1).all() and (w[4] == 0).all() and (w[6] == 0).all()
    assert np.abs(w[5]).max() <= np.sqrt(6.0 / 15)
    c.use_custom_init = False
    for name in ('GlorotUniform', 'GlorotNormal'):
        c.initializer = name
        k = Initializer(c).initial_weights(specs)[0]
        assert np.isfinite(k).all() and 0 < k.std() < 1


def test_gru_initial_weights_follow_keras_defaults(fresh_flags):
    """Keras GRU (reset_after=True): kernel by the configured initializer, orthogonal recurrent kernel, zero [2,3H] bias."""
    c = base_config.get_configs(['--init_scale', '0.25', '--rnn_cell', 'gru'])
    specs = list(zip(orc.param_names(1, 'gru'), orc.param_shapes(1, 8, 3, 12, 'gru')))
    w = Initializer(c).initial_weights(specs)
    assert w[0].shape == (8, 36) and np.abs(w[0]).max() <= 0.25
    np.testing.assert_allclose(w[1] @ w[1].T, np.eye(12), atol=1e-5)
    assert w[2].shape == (2, 36) and (w[2] == 0).all()


def test_model_factory_knows_both_recurrent_families_and_checks_the_uq_flag(fresh_flags):
    """model_utils/model.py:25-33: nn_type is resolved by name; unknown names raise RuntimeError, families that are
    not built NotImplementedError.  (Constructing a model needs a GPU; resolution and refusal do not.)"""
    from lfm_quant_b200.scripts.model_utils import model as factory
    assert {'RNNPointEstimate', 'RNNUqRangeEstimate'} <= set(vars(factory))
    c = base_config.get_configs(['--nn_type', 'MLPPointEstimate'])
    with pytest.raises(NotImplementedError):
        factory.Model(c, None).get_model()
    c.nn_type = 'NoSuchModel'
    with pytest.raises(RuntimeError):
        factory.Model(c, None).get_model()


def test_uq_execution_routing(fresh_flags, monkeypatch):
    """runtime/model_execution.py:149-150,201-206: UQ with training or one process is a single execution; the
    multi-process MC ensemble at predict time is refused here."""
    from lfm_quant_b200.scripts.runtime.model_execution import ModelExecution
    calls = []
    monkeypatch.setattr(ModelExecution, 'single_execution', staticmethod(lambda cfg: calls.append(cfg.train) or 'ran'))
    c = base_config.get_configs(['--nn_type', 'RNNUqRangeEstimate', '--UQ=True', '--train=True', '--num_procs', '4'])
    assert ModelExecution(c)() == 'ran'                  # training: num_procs does not matter
    c.train = False
    with pytest.raises(NotImplementedError):
        ModelExecution(c)()                              # ensemble prediction
    c.num_procs = 1
    assert ModelExecution(c)() == 'ran' and calls == [True, False]


def test_uq_loss_shim_asserts_like_the_reference(fresh_flags):
    """losses.py:145-157: UQ must be on, arguments must be lists."""
    from lfm_quant_b200.scripts.model_utils.losses import Losses
    c = base_config.get_configs(['--nn_type', 'RNNUqRangeEstimate'])
    with pytest.raises(AssertionError, match='UQ should be True'):
        Losses(c, 0).weight_adjusted_uq_loss([0], [0], [0])
    c.UQ = True
    with pytest.raises(AssertionError, match='need to be a list'):
        Losses(c, 0).weight_adjusted_uq_loss(np.zeros(3), [0], [0])


@pytest.mark.parametrize('case', ['defaults', 'overrides', 'bool_forms', 'years'])
def test_flag_parser_matches_the_reference_parser_golden(fresh_flags, case):
    """tests/golden/reference_flags.json holds what the UNMODIFIED reference parser (scripts/configs.py +
    scripts/base_config.py, run by tests/golden/make_reference_flags.py) returns for each argv; this package's mirror
    must return the same value for every one of the reference's flags (it may define additional ones).  The reference's
    CLI flavour (scripts/lfm_quant.py:108-129: dash-separated lists, piecewise flags parsed) cannot be run here -- that
    module imports TensorFlow transitively -- and is covered by test_flag_defaults_match_reference_schema."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_flags.json')))[case]
    assert 'values' in g, g
    c = base_config.get_configs(list(g['argv']), list_sep=',')      # the base_config.py flavour, which is what ran
    diffs = {}
    for k, want in g['values'].items():
        got = getattr(c, k, '<missing>')
        if isinstance(got, tuple):
            got = list(got)
        if got != want:
            diffs[k] = (got, want)
    assert not diffs, diffs
