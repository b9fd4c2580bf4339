"""Replay of the model-facing calls of the reference's two unchanged callers (test plan item 4):

* vae_assoc_ujichar_img_jnt.py  :18-19 seeds, :36 construct_datasets, :53-71 arch dicts, :94-96 train(),
                                :98 save_model(), :112-117 reconstruct / generate / transform on a test batch
* vae_assoc_model_viewer.py     :162-163 (via baxter_vae_assoc_writer.py:94-139) construction with tf.nn.relu +
                                restore_model(folder, fname), :107-113 generate(z_mu) with only row 0 filled

Python-3 transcription of those call sites, textually the same apart from `print` and the data source (synthetic pairs
instead of the un-downloadable pickles), with `vae_assoc_b200.tf_shim` standing in for the `tf` tokens they touch.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import synth                                  # noqa: E402  (test data only)
from oracle import vae_assoc_oracle as vo                 # noqa: E402


def test_ujichar_script_and_model_viewer_call_patterns(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from vae_assoc_b200 import build, dataset, vae_assoc
    from vae_assoc_b200 import tf_shim as tf
    build.build(verbose=False)

    # ---- vae_assoc_ujichar_img_jnt.py -------------------------------------------------------------------------------
    np.random.seed(0)
    tf.set_random_seed(0)
    img_data, fa_data_normed = synth.synth_batch(vo.reference_archs(4), [True, False], 0, 1, 0, 400)
    aug_data = np.concatenate((img_data, fa_data_normed), axis=1).astype(np.float32)
    data_sets = dataset.construct_datasets(aug_data, validation_ratio=.1, test_ratio=.1)
    batch_size, n_z, assoc_lambda, weights = 64, 4, 8, [50, 1]
    img_network_architecture = \
        dict(scope='image', hidden_conv=False, n_hidden_recog_1=500, n_hidden_recog_2=500, n_hidden_gener_1=500,
             n_hidden_gener_2=500, n_input=784, n_z=n_z)
    jnt_network_architecture = \
        dict(scope='joint', hidden_conv=False, n_hidden_recog_1=200, n_hidden_recog_2=200, n_hidden_gener_1=200,
             n_hidden_gener_2=200, n_input=147, n_z=n_z)
    tf.reset_default_graph()
    vae_assoc_model, cost_hist = vae_assoc.train(data_sets, [img_network_architecture, jnt_network_architecture],
                                                 binary=[True, False], weights=weights, assoc_lambda=assoc_lambda,
                                                 learning_rate=0.001, batch_size=batch_size, training_epochs=4,
                                                 display_step=2)
    assert len(cost_hist) == 4 * (data_sets.train._data.shape[0] // batch_size) and np.isfinite(cost_hist).all()
    os.makedirs(str(tmp_path / "output"))
    fname = str(tmp_path / "output" / 'model_batchsize{}_nz{}_lambda{}_weight{}.ckpt'.format(batch_size, n_z, assoc_lambda, weights[0]))
    vae_assoc_model.save_model(fname)
    assert os.path.exists(fname)

    x_sample = data_sets.test.next_batch(batch_size)[0] if data_sets.test._data.shape[0] >= batch_size \
        else data_sets.train.next_batch(batch_size)[0]
    x_sample_seg = [x_sample[:, :784], x_sample[:, 784:]]
    x_reconstruct = vae_assoc_model.reconstruct(x_sample_seg)
    x_synthesis = vae_assoc_model.generate()
    z_test = vae_assoc_model.transform(x_sample_seg)
    for out in (x_reconstruct, x_synthesis):
        assert isinstance(out, list) and [o.shape for o in out] == [(batch_size, 784), (batch_size, 147)]
        assert all(o.dtype == np.float32 and np.isfinite(o).all() for o in out)
        assert out[0].min() >= 0.0 and out[0].max() <= 1.0                  # Bernoulli means (sigmoid)
    assert [z.shape for z in z_test] == [(batch_size, n_z)] * 2
    # the script prints the first five codes of both modalities side by side (:185-189): association pulls them together
    gap = np.abs(z_test[0] - z_test[1]).mean()
    spread = np.abs(z_test[0] - z_test[0].mean(axis=0)).mean()
    assert np.isfinite(gap) and np.isfinite(spread)
    want_row0 = vae_assoc_model.generate(z_mu=np.tile(z_test[0][:1], (batch_size, 1)))

    # ---- vae_assoc_model_viewer.py / baxter_vae_assoc_writer.py:94-139 -----------------------------------------------
    tf.reset_default_graph()
    viewer_model = vae_assoc.AssocVariationalAutoEncoder([img_network_architecture, jnt_network_architecture],
                                                         binary=[True, False], transfer_fct=tf.nn.relu,
                                                         assoc_lambda=5, learning_rate=0.0001, batch_size=batch_size)
    assert len(tf.all_variables()) == 86                    # baxter_vae_assoc_writer.py:599 prints this count
    folder, name = os.path.split(fname)
    viewer_model.restore_model(folder, name)
    z_mu = np.zeros((viewer_model.batch_size, n_z))
    z_mu[0, :] = z_test[0][0]                                # only row 0 carries the sliders' values (:108-111)
    x_reconstr_means = viewer_model.generate(z_mu=z_mu)
    img = np.reshape(x_reconstr_means[0][0], (28, 28))
    fa_parms = x_reconstr_means[1][0]
    assert img.shape == (28, 28) and fa_parms.shape == (147,)
    # the restored model decodes row 0 exactly like the trained one
    np.testing.assert_allclose(x_reconstr_means[0][0], want_row0[0][0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(x_reconstr_means[1][0], want_row0[1][0], rtol=1e-5, atol=1e-6)
    viewer_model.close()
