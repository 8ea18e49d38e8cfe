"""Batching server behind the reference's pipe protocol.

Same surface as ``ChessModelAPI`` (/root/reference/src/chess_zero/agent/api_chess.py:14-69):
``ChessModelAPI(agent_model)``, ``.start()``, ``.create_pipe()``; clients ``send`` one float32 (18,8,8)
observation and ``recv`` ``(policy float32[1968], float value)`` (agent/player_chess.py:237-248), one
outstanding request per pipe.  Whatever is ready when the poll returns forms the batch, replies go back
in request order.

Differences from the reference, all on the failure side (it has none): a failing forward
pass is delivered to every waiting client as the exception object instead of leaving them blocked in
``recv()`` forever; closed pipes are dropped; ``stop()`` exists.
"""
import pickle
import selectors
import time
from multiprocessing import Pipe
from threading import Lock, Thread

import numpy as np

from . import shm_pipes


class ChessModelAPI:
    def __init__(self, agent_model, transport="pipe", shm_capacity=1024, planes_shape=(18, 8, 8), n_labels=1968):
        if transport not in ("pipe", "shm"):
            raise ValueError(f"transport must be 'pipe' or 'shm', got {transport!r}")
        self.agent_model = agent_model
        self.transport = transport
        self.pipes = []
        self._lock = Lock()
        self._stop = False
        self._thread = None
        self._selector = selectors.DefaultSelector()
        self._slab = shm_pipes.Slab(shm_capacity, planes_shape, n_labels) if transport == "shm" else None
        self.batches = 0
        self.positions = 0

    def start(self):
        self._thread = Thread(target=self._predict_batch_worker, name="prediction_worker", daemon=True)
        self._thread.start()

    def stop(self):
        self._stop = True
        if self._thread is not None:
            self._thread.join(timeout=2.0)
        if self._slab is not None:
            self._slab.close()
            self._slab = None

    def create_pipe(self):
        with self._lock:
            if self._slab is not None:
                me, slot, you = self._slab.endpoint()
            else:
                (me, you), slot = Pipe(), None
            self.pipes.append(me)
            self._selector.register(me, selectors.EVENT_READ, slot)
        return you

    def _drop(self, pipe):
        with self._lock:
            if pipe in self.pipes:
                self.pipes.remove(pipe)
                try:
                    self._selector.unregister(pipe)
                except (KeyError, ValueError, OSError):
                    pass
        try:
            pipe.close()
        except OSError:
            pass

    def _gather(self, ready):
        """ready: [(server-side Connection, slot)] -> (batch, pipes to answer, their slots)."""
        result_pipes, slots, data = [], [], []
        for pipe, slot in ready:
            try:
                if slot is None:
                    while pipe.poll():
                        data.append(pipe.recv())
                        result_pipes.append(pipe)
                else:
                    pipe.recv_bytes()
                    slots.append(slot)
                    result_pipes.append(pipe)
            except (EOFError, OSError):
                self._drop(pipe)
        if not result_pipes:
            return None, result_pipes, slots
        if self._slab is not None:
            return self._slab.planes[slots], result_pipes, slots       # one gather copy, already float32
        return np.asarray(data, dtype=np.float32), result_pipes, slots

    def _predict_batch_worker(self):
        while not self._stop:
            with self._lock:
                events = self._selector.select(timeout=0.001) if self.pipes else []
            if not events:
                if not self.pipes:
                    time.sleep(0.001)
                continue
            batch, result_pipes, slots = self._gather([(key.fileobj, key.data) for key, _ in events])
            if batch is None:
                continue
            try:
                policy_ary, value_ary = self.agent_model.model.predict_on_batch(batch)
                value_ary = np.asarray(value_ary).reshape(-1)
                if self._slab is not None:
                    self._slab.policy[slots] = policy_ary
                    self._slab.value[slots] = value_ary
                else:
                    replies = [(p, float(v)) for p, v in zip(policy_ary, value_ary)]
            except Exception as exc:  # noqa: BLE001 - delivered to the clients, then re-raised
                for pipe in result_pipes:
                    try:
                        if self._slab is not None:
                            pipe.send_bytes(shm_pipes.ERR + pickle.dumps(exc))
                        else:
                            pipe.send(exc)
                    except (OSError, ValueError):
                        pass
                raise
            self.batches += 1
            self.positions += len(result_pipes)
            for i, pipe in enumerate(result_pipes):
                try:
                    if self._slab is not None:
                        pipe.send_bytes(shm_pipes.OK)
                    else:
                        pipe.send(replies[i])
                except (OSError, ValueError):
                    self._drop(pipe)
