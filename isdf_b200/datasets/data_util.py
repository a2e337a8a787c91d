"""Keyframe replay buffer -- mirror of reference isdf/datasets/data_util.py:11-102 (FrameData).

Same fields and the same append / replace-last semantics; device tensors grow geometrically in
place of a torch.cat per keyframe (the reference re-allocates the whole buffer on every add)."""
import numpy as np
import torch

_FIELDS = ("frame_id", "im_batch", "im_batch_np", "depth_batch", "depth_batch_np", "T_WC_batch", "T_WC_batch_np",
           "normal_batch", "frame_avg_losses", "T_WC_track", "T_WC_gt")


class _Growable:
    """Rows appended into a capacity-doubling buffer; `view` is the live [n, ...] prefix."""

    def __init__(self, wrap=None):
        self.buf, self.n = None, 0
        if wrap is not None:                 # adopt the caller's rows as-is (no copy, no spare capacity)
            self.buf, self.n = wrap, wrap.shape[0]

    def append(self, rows):
        k = rows.shape[0]
        if self.buf is None:
            cap = max(8, k)
            self.buf = (np.empty((cap,) + rows.shape[1:], rows.dtype) if isinstance(rows, np.ndarray)
                        else torch.empty((cap,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device))
        if self.n + k > self.buf.shape[0]:
            cap = max(2 * self.buf.shape[0], self.n + k)
            if isinstance(self.buf, np.ndarray):
                new = np.empty((cap,) + self.buf.shape[1:], self.buf.dtype)
            else:
                new = torch.empty((cap,) + tuple(self.buf.shape[1:]), dtype=self.buf.dtype, device=self.buf.device)
            new[:self.n] = self.buf[:self.n]
            self.buf = new
        self.buf[self.n:self.n + k] = rows
        self.n += k

    def replace_last(self, rows):
        self.buf[self.n - 1] = rows[0]

    @property
    def view(self):
        return None if self.buf is None else self.buf[:self.n]


class _LazyRows:
    """Host (numpy) rows kept by reference; the [n, ...] batch is materialised only when somebody reads it
    (the drivers read im_batch_np / depth_batch_np / T_WC_batch_np for visualisation, never on the step)."""

    def __init__(self, wrap=None):
        self.rows = [] if wrap is None else [wrap]
        self._cat = wrap

    def append(self, rows):
        self.rows.append(rows)
        self._cat = None

    def replace_last(self, rows):
        last = self.rows[-1]
        if last.shape[0] == 1:
            self.rows[-1] = rows[:1]
        else:
            self.rows[-1] = last[:-1]
            self.rows.append(rows[:1])
        self._cat = None

    @property
    def n(self):
        return sum(r.shape[0] for r in self.rows)

    @property
    def view(self):
        if self._cat is None:
            self._cat = self.rows[0] if len(self.rows) == 1 else np.concatenate(self.rows)
            self.rows = [self._cat]
        return self._cat


def _holder(rows):
    big = isinstance(rows, np.ndarray) and rows.ndim >= 3
    return _LazyRows(wrap=rows) if big else _Growable(wrap=rows)


class FrameData:
    def __init__(self, frame_id=None, im_batch=None, im_batch_np=None, depth_batch=None, depth_batch_np=None,
                 T_WC_batch=None, T_WC_batch_np=None, normal_batch=None, frame_avg_losses=None, T_WC_track=None,
                 T_WC_gt=None):
        self._store = {}
        for name, val in zip(_FIELDS, (frame_id, im_batch, im_batch_np, depth_batch, depth_batch_np, T_WC_batch,
                                       T_WC_batch_np, normal_batch, frame_avg_losses, T_WC_track, T_WC_gt)):
            if val is not None:
                setattr(self, name, val)
        self.count = 0 if frame_id is None else len(frame_id)

    def __getattr__(self, name):
        if name in _FIELDS:
            g = self.__dict__.get("_store", {}).get(name)
            return None if g is None else g.view
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in _FIELDS:
            if value is not None:
                self._store[name] = _holder(value)
            else:
                self._store.pop(name, None)
        else:
            object.__setattr__(self, name, value)

    def _add(self, name, rows, replace):
        if rows is None:
            return
        g = self._store.get(name)
        if g is None:       # first rows are adopted, not copied (the reference assigns them too, data_util.py:52-60)
            self._store[name] = _holder(rows)
        elif replace:
            g.replace_last(rows)
        else:
            g.append(rows)

    def add_frame_data(self, data, replace):
        """Append `data` (a FrameData) or overwrite the last row (data_util.py:45-78)."""
        for name in ("frame_id", "im_batch", "im_batch_np", "depth_batch", "depth_batch_np", "T_WC_batch",
                     "T_WC_batch_np", "normal_batch"):
            self._add(name, getattr(data, name), replace)
        ref = data.depth_batch if data.depth_batch is not None else data.im_batch
        zeros = torch.zeros([ref.shape[0]], device=ref.device)
        self._add("frame_avg_losses", zeros, replace)
        if data.T_WC_gt is not None:
            self._add("T_WC_gt", data.T_WC_gt, replace)

    def __len__(self):
        fid = self.frame_id
        return 0 if fid is None else len(fid)


def expand_data(batch, data, replace=False):
    """Concatenate `data` under `batch`, or overwrite the last row (data_util.py:84-102)."""
    if batch is None:
        return data
    if replace:
        batch[-1] = data[0]
        return batch
    return torch.cat((batch, data)) if torch.is_tensor(data) else np.concatenate((batch, data))
