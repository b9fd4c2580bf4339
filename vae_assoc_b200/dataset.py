"""Host-side batching with the reference's observable behaviour (/root/reference/dataset.py:6-72): in-RAM rows, one
shuffle at construction, train / validation / test split by ratio, sequential `next_batch` that starts a new epoch --
with a fresh shuffle -- when a batch would run past the end.

The rows are never physically re-ordered here: a data set keeps ONE immutable row block plus an index vector, and an
epoch shuffle composes the new permutation into that vector (the reference re-gathers the whole array every epoch,
dataset.py:27-33).  The numpy RNG is consumed exactly as the reference consumes it (one `np.random.shuffle` of
`arange(n)` per shuffle), so a caller that seeds `np.random` (vae_assoc_ujichar_img_jnt.py:18) sees the same batches.
Pure host bookkeeping: no arithmetic of the hot path lives here; benchmarks use the device-resident synthetic generator
instead (AssocVariationalAutoEncoder.synth_batch).
"""
import numpy as np


class DataSets(object):
    """Attribute bag: `.train`, `.validation`, `.test` (dataset.py:3-4)."""
    pass


def _draw_permutation(n):
    # the reference's RNG call pattern: arange + in-place shuffle (dataset.py:29-30, :50-51)
    perm = np.arange(n)
    np.random.shuffle(perm)
    return perm


class DataSet(object):
    def __init__(self, data, labels=None):
        if labels is not None and data.shape[0] != labels.shape[0]:
            raise AssertionError('data.shape: %s labels.shape: %s' % (data.shape, labels.shape))
        self._rows = data
        self._row_labels = labels
        self._order = None                  # None = identity; else current epoch's row order
        self._cursor = 0
        self._epochs_completed = 0

    # names the reference's callers read (vae_assoc.py:508,528: `_data.shape[0]`)
    @property
    def _num_examples(self):
        return self._rows.shape[0]

    @property
    def _index_in_epoch(self):
        return self._cursor

    @property
    def _data(self):
        return self._rows if self._order is None else self._rows[self._order]

    @property
    def _labels(self):
        if self._row_labels is None or self._order is None:
            return self._row_labels
        return self._row_labels[self._order]

    def _take(self, lo, hi):
        if self._order is None:
            rows = self._rows[lo:hi]
            labs = None if self._row_labels is None else self._row_labels[lo:hi]
        else:
            idx = self._order[lo:hi]
            rows = self._rows[idx]
            labs = None if self._row_labels is None else self._row_labels[idx]
        return rows, labs

    def _advance(self, batch_size):
        n = self._num_examples
        lo, hi = self._cursor, self._cursor + batch_size
        if hi > n:
            assert batch_size <= n
            self._epochs_completed += 1
            perm = _draw_permutation(n)
            self._order = perm if self._order is None else self._order[perm]
            lo, hi = 0, batch_size
        self._cursor = hi
        return lo, hi

    def next_batch(self, batch_size):
        """The next `batch_size` examples; a batch that would cross the end starts the next epoch instead, on a freshly
        shuffled order (the tail of the old epoch is dropped, as in dataset.py:25-38)."""
        lo, hi = self._advance(batch_size)
        return self._take(lo, hi)

    def next_indices(self, batch_size):
        """Row indices (into the immutable row block `_rows`) of the batch `next_batch` would return, with the same
        cursor / epoch / RNG behaviour -- for the device-resident data-set path, which gathers the rows on the GPU."""
        lo, hi = self._advance(batch_size)
        if self._order is None:
            return np.arange(lo, hi, dtype=np.int64)
        return np.ascontiguousarray(self._order[lo:hi], dtype=np.int64)


def construct_datasets(data, labels=None, shuffle=True, validation_ratio=.1, test_ratio=.1):
    """Split `data` (and `labels`) into train / validation / test by the given ratios after one optional shuffle
    (dataset.py:45-72: test = last `test_ratio`, validation = the `validation_ratio` before it)."""
    n = data.shape[0]
    if labels is not None and labels.shape[0] != n:
        raise AssertionError('data.shape: %s labels.shape: %s' % (data.shape, labels.shape))
    if shuffle:
        perm = _draw_permutation(n)
        data = data[perm]
        labels = None if labels is None else labels[perm]
    cut_test = int((1 - test_ratio) * n)
    cut_valid = int((1 - validation_ratio - test_ratio) * n)
    bounds = {"train": (0, cut_valid), "validation": (cut_valid, cut_test), "test": (cut_test, n)}
    out = DataSets()
    for name, (a, b) in bounds.items():
        setattr(out, name, DataSet(data[a:b], None if labels is None else labels[a:b]))
    return out
