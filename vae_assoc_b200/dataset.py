"""Host-side batching with the reference's semantics (dataset.py:6-72): in-RAM rows, shuffle, train / validation /
test split, sequential `next_batch` with a reshuffle at each epoch wrap.  Pure host bookkeeping (no arithmetic of
the hot path lives here); benchmarks use the device-resident synthetic generator instead
(AssocVariationalAutoEncoder.synth_batch)."""
import numpy as np


class DataSets(object):
    pass


class DataSet(object):
    def __init__(self, data, labels=None):
        if labels is not None:
            assert data.shape[0] == labels.shape[0], (
                'data.shape: %s labels.shape: %s' % (data.shape, labels.shape))
        self._num_examples = data.shape[0]
        self._data = data
        self._labels = labels
        self._epochs_completed = 0
        self._index_in_epoch = 0

    def next_batch(self, batch_size):
        """Return the next `batch_size` examples from this data set (dataset.py:22-43)."""
        start = self._index_in_epoch
        self._index_in_epoch += batch_size
        if self._index_in_epoch > self._num_examples:
            self._epochs_completed += 1
            perm = np.arange(self._num_examples)
            np.random.shuffle(perm)
            self._data = self._data[perm]
            if self._labels is not None:
                self._labels = self._labels[perm]
            start = 0
            self._index_in_epoch = batch_size
            assert batch_size <= self._num_examples
        end = self._index_in_epoch
        if self._labels is not None:
            return self._data[start:end], self._labels[start:end]
        return self._data[start:end], None


def construct_datasets(data, labels=None, shuffle=True, validation_ratio=.1, test_ratio=.1):
    """dataset.py:45-72"""
    data_sets = DataSets()
    if shuffle:
        perm = np.arange(data.shape[0])
        np.random.shuffle(perm)
        data = data[perm]
        if labels is not None:
            labels = labels[perm]
    n = data.shape[0]
    test_start = int((1 - test_ratio) * n)
    valid_start = int((1 - validation_ratio - test_ratio) * n)
    lab = (lambda a, b: labels[a:b]) if labels is not None else (lambda a, b: None)
    data_sets.train = DataSet(data[:valid_start], lab(0, valid_start))
    data_sets.validation = DataSet(data[valid_start:test_start], lab(valid_start, test_start))
    data_sets.test = DataSet(data[test_start:], lab(test_start, n))
    return data_sets
