"""Host-side permutation stream for the randomization null.

The reference draws one `np.random.permutation(indx_vals)` per iteration from the legacy global MT19937 stream
after `np.random.seed(random_seed)` and applies it IN PLACE to the already permuted attribute matrix
(safepy/safe_extras.py:46-58).  The shuffle is inherently sequential, so it stays on the host; what goes to the GPU
is the composition of those shuffles as plain gather indices: rows[p, t] is the row of the ORIGINAL matrix that node
t holds during permutation p.  Rows without any data are never moved (safe_extras.py:51).
"""
import numpy as np


def rows_with_data(node2attribute):
    """indx_vals of safe_extras.py:51: the rows with at least one non-NaN entry, ascending.

    Upstream evaluates np.sum(~np.isnan(B), axis=1) over the whole matrix (1 s for 100 000 x 5000 float32).  A row is
    settled by its first non-NaN entry, so a few leading columns settle almost every row and only the rest is scanned
    in full."""
    a = np.asarray(node2attribute)
    if a.ndim != 2:
        return np.nonzero(np.sum(~np.isnan(a), axis=1))[0]
    has = np.zeros(a.shape[0], dtype=bool)
    open_rows = np.arange(a.shape[0])
    for j in range(min(4, a.shape[1])):
        if open_rows.size * 8 < a.shape[0]:
            break
        found = ~np.isnan(a[open_rows, j])
        has[open_rows[found]] = True
        open_rows = open_rows[~found]
    if open_rows.size:
        has[open_rows] = ~np.isnan(a[open_rows]).all(axis=1)
    return np.nonzero(has)[0]


def native_seed(random_seed):
    """True when the library's native replay (sb_perm_stream_*) covers this seed: None or an int in [0, 2**32) --
    what SAFE.random_seed holds (safe.py:102, 180-184).  Other seeds (arrays, ...) take NumPy's own seeding path."""
    if random_seed is None:
        return True
    return isinstance(random_seed, (int, np.integer)) and not isinstance(random_seed, bool) \
        and 0 <= int(random_seed) < 2 ** 32


def perm_stream(node2attribute, random_seed):
    """The library's host replay of the reference's RNG calls for this attribute matrix (see _lib.PermStream)."""
    from . import _lib
    return _lib.PermStream(node2attribute.shape[0], rows_with_data(node2attribute), random_seed)


def make_perm_rows(node2attribute, num_permutations, random_seed, out=None):
    """Replay the reference's RNG calls and compose them. Returns int32 [num_permutations, n].

    Leaves the global NumPy RNG exactly where the reference leaves it, so interleaving with reference code keeps both
    streams aligned.  random_seed=None seeds from OS entropy (np.random.seed(None)), as upstream."""
    n = node2attribute.shape[0]
    rows = out if out is not None else np.empty((num_permutations, n), dtype=np.int32)
    if native_seed(random_seed):
        stream = perm_stream(node2attribute, random_seed)
        stream.next(num_permutations, rows)
        stream.sync_numpy()
        stream.close()
        return rows
    np.random.seed(random_seed)
    indx_vals = rows_with_data(node2attribute)
    cur = np.arange(n, dtype=np.int32)
    for p in range(num_permutations):
        # n2a[indx_vals, :] = n2a[np.random.permutation(indx_vals), :]
        cur[indx_vals] = cur[np.random.permutation(indx_vals)]
        rows[p] = cur
    return rows


class iter_perm_rows:
    """The same stream as make_perm_rows, produced piece by piece by a background thread that starts at once, so that
    the GPU counts piece k (the C call releases the GIL) while the host replays the RNG for piece k + 1.  Pieces
    start small (the device gets work after a few permutations) and double up to `piece`.  Iterating yields int32
    [<= piece, n] arrays that together equal make_perm_rows(...); close() (or exhausting it) stops the thread."""

    def __init__(self, node2attribute, num_permutations, random_seed, piece=128, depth=4, first=16):
        import queue
        import threading
        self._queue = queue.Queue(maxsize=depth)
        self._stop = threading.Event()
        self._done = False
        n = node2attribute.shape[0]
        indx_vals = rows_with_data(node2attribute)
        piece = max(1, int(piece))

        def put(item):
            while not self._stop.is_set():
                try:
                    self._queue.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def produce():
            try:
                np.random.seed(random_seed)
                cur = np.arange(n, dtype=np.int32)
                p0, size = 0, max(1, min(int(first), piece))
                while p0 < num_permutations:
                    rows = np.empty((min(size, num_permutations - p0), n), dtype=np.int32)
                    for k in range(rows.shape[0]):
                        cur[indx_vals] = cur[np.random.permutation(indx_vals)]
                        rows[k] = cur
                    if not put(rows):
                        return
                    p0 += rows.shape[0]
                    size = min(piece, size * 2)
                put(None)
            except BaseException as exc:  # noqa: BLE001  (handed to the consumer)
                put(exc)

        self._worker = threading.Thread(target=produce, name="safe-b200-perm-replay", daemon=True)
        self._worker.start()

    def __iter__(self):
        return self

    def __next__(self):
        if self._done:
            raise StopIteration
        item = self._queue.get()
        if item is None or isinstance(item, BaseException):
            self.close()
            if item is None:
                raise StopIteration
            raise item
        return item

    def close(self):
        self._done = True
        self._stop.set()
        self._worker.join()


def shard_bounds(num_permutations, world_size, rank):
    """Contiguous permutation range [lo, hi) owned by `rank` (the last ranks get the short shards)."""
    per = -(-num_permutations // world_size)
    lo = min(num_permutations, rank * per)
    return lo, min(num_permutations, lo + per)
