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
from multiprocessing import Pipe, connection
from threading import Lock, Thread

import numpy as np


class ChessModelAPI:
    def __init__(self, agent_model):
        self.agent_model = agent_model
        self.pipes = []
        self._lock = Lock()
        self._stop = False
        self._thread = None
        self.batches = 0
        self.positions = 0

    def start(self):
        self._thread = Thread(target=self._predict_batch_worker, name="prediction_worker", daemon=True)
        self._thread.start()

    def stop(self):
        self._stop = True
        if self._thread is not None:
            self._thread.join(timeout=2.0)

    def create_pipe(self):
        me, you = Pipe()
        with self._lock:
            self.pipes.append(me)
        return you

    def _drop(self, pipe):
        with self._lock:
            if pipe in self.pipes:
                self.pipes.remove(pipe)
        try:
            pipe.close()
        except OSError:
            pass

    def _predict_batch_worker(self):
        while not self._stop:
            with self._lock:
                pipes = list(self.pipes)      # create_pipe may append while we wait
            if not pipes:
                connection.wait([], timeout=0.001)
                continue
            ready = connection.wait(pipes, timeout=0.001)
            if not ready:
                continue
            data, result_pipes = [], []
            for pipe in ready:
                try:
                    while pipe.poll():
                        data.append(pipe.recv())
                        result_pipes.append(pipe)
                except (EOFError, OSError):
                    self._drop(pipe)
            if not data:
                continue
            try:
                batch = np.asarray(data, dtype=np.float32)
                policy_ary, value_ary = self.agent_model.model.predict_on_batch(batch)
                replies = [(p, float(v)) for p, v in zip(policy_ary, np.asarray(value_ary).reshape(-1))]
            except Exception as exc:  # noqa: BLE001 - delivered to the clients, then re-raised
                for pipe in result_pipes:
                    try:
                        pipe.send(exc)
                    except (OSError, ValueError):
                        pass
                raise
            self.batches += 1
            self.positions += len(replies)
            for pipe, reply in zip(result_pipes, replies):
                try:
                    pipe.send(reply)
                except (OSError, ValueError):
                    self._drop(pipe)
