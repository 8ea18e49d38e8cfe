"""Shared-memory transport with the reference's pipe contract.

The reference moves every request and reply through ``multiprocessing.Pipe`` (agent/api_chess.py:40-41,
agent/player_chess.py:237-248): 4.6 KB of planes pickled one way, 7.9 KB of policy pickled back, per leaf.  Here the
payload lives in one shared slab, a slot per pipe, and the OS pipe carries a single wake-up byte each way:

    client.send(planes)  ->  copy into slot.planes, 1 byte on the pipe
    server               ->  gathers the ready slots into one batch, runs it, scatters policy/value into the slots,
                             1 byte back per client
    client.recv()        ->  blocks on the byte, returns (policy copy, float value)

The endpoint object has the ``Connection`` methods the reference's clients use (``send``, ``recv``, ``poll``,
``close``) and pickles across processes the way a ``Connection`` does (through multiprocessing's own pickler: a
``Process`` argument, a ``Manager().list`` entry as in worker/self_play.py:47), attaching to the slab by name on
first use.  One outstanding request per endpoint, as in the reference.
"""
import pickle
from multiprocessing import Pipe, resource_tracker, shared_memory

import numpy as np

WAKE = b"r"
OK = b"k"
ERR = b"e"


def slab_layout(n_slots, planes_shape, n_labels):
    planes_elems = int(np.prod(planes_shape))
    off_policy = n_slots * planes_elems * 4
    off_value = off_policy + n_slots * n_labels * 4
    return off_policy, off_value, off_value + n_slots * 4


def slab_views(buf, n_slots, planes_shape, n_labels):
    off_policy, off_value, _ = slab_layout(n_slots, planes_shape, n_labels)
    planes = np.ndarray((n_slots,) + tuple(planes_shape), np.float32, buffer=buf)
    policy = np.ndarray((n_slots, n_labels), np.float32, buffer=buf, offset=off_policy)
    value = np.ndarray((n_slots,), np.float32, buffer=buf, offset=off_value)
    return planes, policy, value


class Slab:
    """Server side: owns the shared segment; ``capacity`` endpoints can be handed out."""

    def __init__(self, capacity, planes_shape=(18, 8, 8), n_labels=1968):
        self.capacity, self.planes_shape, self.n_labels = int(capacity), tuple(planes_shape), int(n_labels)
        size = slab_layout(self.capacity, self.planes_shape, self.n_labels)[2]
        self.shm = shared_memory.SharedMemory(create=True, size=size)
        self.planes, self.policy, self.value = slab_views(self.shm.buf, self.capacity, self.planes_shape, self.n_labels)
        self.used = 0

    def endpoint(self):
        """Returns (server-side Connection, slot, client ShmPipe)."""
        if self.used >= self.capacity:
            raise RuntimeError(f"shared slab holds {self.capacity} pipes; raise shm_capacity")
        slot = self.used
        self.used += 1
        mine, theirs = Pipe()
        return mine, slot, ShmPipe(self.shm.name, slot, self.capacity, self.planes_shape, self.n_labels, theirs)

    def close(self):
        self.planes = self.policy = self.value = None
        try:
            self.shm.close()
            self.shm.unlink()
        except (FileNotFoundError, BufferError):
            pass


class ShmPipe:
    """Client end: what ``ChessPlayer.predict`` holds instead of a ``Connection``."""

    def __init__(self, name, slot, n_slots, planes_shape, n_labels, conn):
        self._name, self._slot, self._n_slots = name, slot, n_slots
        self._planes_shape, self._n_labels, self._conn = tuple(planes_shape), n_labels, conn
        self._shm = None

    def __reduce__(self):
        return ShmPipe, (self._name, self._slot, self._n_slots, self._planes_shape, self._n_labels, self._conn)

    def _attach(self):
        self._shm = shared_memory.SharedMemory(name=self._name)
        try:        # the creator owns the segment; keep this process's exit from unlinking it (bpo-39959)
            resource_tracker.unregister(self._shm._name, "shared_memory")
        except Exception:  # noqa: BLE001
            pass
        planes, policy, value = slab_views(self._shm.buf, self._n_slots, self._planes_shape, self._n_labels)
        self._planes, self._policy, self._value = planes[self._slot], policy[self._slot], value[self._slot:self._slot + 1]

    def send(self, planes):
        if self._shm is None:
            self._attach()
        np.copyto(self._planes, planes, casting="same_kind")
        self._conn.send_bytes(WAKE)

    def recv(self):
        msg = self._conn.recv_bytes()
        if msg[:1] == ERR:
            return pickle.loads(msg[1:])
        return self._policy.copy(), float(self._value[0])

    def poll(self, timeout=0.0):
        return self._conn.poll(timeout)

    def fileno(self):
        return self._conn.fileno()

    def close(self):
        self._planes = self._policy = self._value = None
        if self._shm is not None:
            try:
                self._shm.close()
            except BufferError:
                pass
            self._shm = None
        self._conn.close()
