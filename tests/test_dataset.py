"""Host batching semantics of vae_assoc_b200/dataset.py against the behaviour documented for the reference
(/root/reference/dataset.py:6-72): one shuffle at construction, split by ratio, sequential batches, a fresh shuffle of
the CURRENT order at every epoch wrap (the tail that does not fill a batch is dropped), numpy RNG consumed identically."""
import numpy as np

from vae_assoc_b200 import dataset


def _expected_stream(data, batch, n_batches, seed, validation_ratio=.1, test_ratio=.1):
    """Plain restatement of the documented behaviour on the TRAIN split, physically re-gathering rows."""
    np.random.seed(seed)
    perm = np.arange(data.shape[0]); np.random.shuffle(perm)
    rows = data[perm]
    n = rows.shape[0]
    train = rows[:int((1 - validation_ratio - test_ratio) * n)]
    out, pos = [], 0
    for _ in range(n_batches):
        if pos + batch > train.shape[0]:
            p = np.arange(train.shape[0]); np.random.shuffle(p)
            train = train[p]; pos = 0
        out.append(train[pos:pos + batch]); pos += batch
    return out, train.shape[0]


def test_split_sizes_and_batches_match_reference_behaviour():
    data = np.arange(103 * 3, dtype=np.float32).reshape(103, 3)
    want, n_train = _expected_stream(data, 16, 23, seed=5)
    np.random.seed(5)
    ds = dataset.construct_datasets(data, validation_ratio=.1, test_ratio=.1)
    assert ds.train._data.shape[0] == n_train == int(0.8 * 103)
    assert ds.validation._data.shape[0] == int(0.9 * 103) - int(0.8 * 103)
    assert ds.test._data.shape[0] == 103 - int(0.9 * 103)
    for k, w in enumerate(want):
        got, labels = ds.train.next_batch(16)
        assert labels is None
        np.testing.assert_array_equal(got, w, err_msg="batch %d" % k)
    assert ds.train._epochs_completed == 23 // (n_train // 16) - (1 if 23 % (n_train // 16) == 0 else 0)
    # every split is disjoint and together they are the data
    allrows = np.concatenate([ds.train._data, ds.validation._data, ds.test._data])
    assert sorted(map(tuple, allrows)) == sorted(map(tuple, data))


def test_labels_travel_with_rows_and_no_shuffle_keeps_order():
    data = np.arange(40, dtype=np.float32).reshape(20, 2)
    labels = np.arange(20).reshape(20, 1)
    ds = dataset.construct_datasets(data, labels, shuffle=False, validation_ratio=.2, test_ratio=.2)
    x, y = ds.train.next_batch(5)
    np.testing.assert_array_equal(x, data[:5]); np.testing.assert_array_equal(y, labels[:5])
    np.random.seed(1)
    for _ in range(7):
        x, y = ds.train.next_batch(5)
        np.testing.assert_array_equal(x[:, 0] // 2, y[:, 0])        # row i carries label i through every reshuffle
    np.testing.assert_array_equal(ds.train._data[:, 0] // 2, ds.train._labels[:, 0])


def test_extract_loaders_follow_reference_order(tmp_path):
    """utils.extract_images / extract_jnt_fa_parms (/root/reference/utils.py:142-195): scaling, ordering, digit filter,
    and the Python-2 pickle the reference's files are (protocol 2, latin1)."""
    import pickle
    from vae_assoc_b200 import utils
    rng = np.random.RandomState(3)
    # keys as in the UJI set: the ordering of images is by the LAST letter, that of the parameters by the whole key
    imgs = {"b": [rng.randint(0, 256, (28, 28)).astype(np.uint8) for _ in range(3)],
            "a": [rng.randint(0, 256, (28, 28)).astype(np.uint8) for _ in range(2)],
            "7": [rng.randint(0, 256, (28, 28)).astype(np.uint8) for _ in range(2)]}
    fa = {k: [rng.normal(size=147) for _ in v] for k, v in imgs.items()}
    x = utils.extract_images(data=imgs, only_digits=False)
    assert x.shape == (7, 784) and x.dtype == np.float32 and 0.0 <= x.min() and x.max() <= 1.0
    np.testing.assert_allclose(x[0], imgs["7"][0].reshape(-1) / 255.0, rtol=1e-6)       # '7' < 'a' < 'b'
    np.testing.assert_allclose(x[2], imgs["a"][0].reshape(-1) / 255.0, rtol=1e-6)
    assert utils.extract_images(data=imgs, only_digits=True).shape == (2, 784)
    p, mean, std = utils.extract_jnt_fa_parms(data=fa, only_digits=False)
    assert p.shape == (7, 147)
    np.testing.assert_allclose(mean, p.mean(0)); np.testing.assert_allclose(std, p.std(0))
    np.testing.assert_array_equal(p[0], fa["7"][0])
    assert utils.extract_jnt_trajs(data={"3": [np.ones((5, 7))]}).shape == (1, 35)
    # from files, as the training script does (vae_assoc_ujichar_img_jnt.py:21-36)
    fi, ff = tmp_path / "img.pkl", tmp_path / "fa.pkl"
    fi.write_bytes(pickle.dumps(imgs, protocol=2)); ff.write_bytes(pickle.dumps(fa, protocol=2))
    np.random.seed(0)
    ds, m2, s2 = utils.load_paired_datasets(str(fi), str(ff))
    assert ds.train._data.shape[1] == 931
    assert ds.train._data.shape[0] + ds.validation._data.shape[0] + ds.test._data.shape[0] == 7
    np.testing.assert_allclose(m2, mean)


def test_next_indices_equals_next_batch():
    from vae_assoc_b200 import dataset
    data = np.arange(53 * 3, dtype=np.float32).reshape(53, 3)
    a, b = dataset.DataSet(data), dataset.DataSet(data)
    np.random.seed(4)                         # the epoch shuffles draw from the global numpy RNG (dataset.py:29-30)
    batches = [a.next_batch(8)[0] for _ in range(17)]     # crosses several epoch boundaries (53 rows, batches of 8)
    np.random.seed(4)
    for rows in batches:
        idx = b.next_indices(8)
        assert idx.dtype == np.int64
        np.testing.assert_array_equal(rows, data[idx])
    assert a._epochs_completed == b._epochs_completed > 0
