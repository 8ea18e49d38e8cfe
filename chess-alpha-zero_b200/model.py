"""``ChessModel`` with the reference's surface, backed by the sm_100a evaluator.

Mirrors /root/reference/src/chess_zero/agent/model_chess.py:26-166 -- ``build``, ``load``, ``save``,
``get_pipes``, ``fetch_digest``, attributes ``model`` / ``digest`` / ``api`` -- so the reference's workers
(self_play.py:43-48, evaluate.py:44,68, uci.py:63-69) and its own, unmodified ``ChessModelAPI`` can run on
it: ``self.model`` is an :class:`Evaluator` whose ``predict_on_batch`` replaces the Keras call
(api_chess.py:67).  The FTP push/pull of the "distributed" mode (model_chess.py:130-141,170-186) is out
of scope.
"""
import hashlib
import json
import os
from logging import getLogger

from . import keras_graph
from .api import ChessModelAPI
from .config import Config
from .evaluator import Evaluator

logger = getLogger(__name__)


class ChessModel:
    def __init__(self, config=None, device=0, precision="bf16", max_batch=4096, transport="pipe"):
        self.config = config if config is not None else Config()
        self.device = device
        self.precision = precision
        self.max_batch = max_batch
        self.transport = transport     # "pipe": multiprocessing.Pipe payloads (the reference's); "shm": shm_pipes.py
        self.model = None      # Evaluator; the attribute name the reference's ChessModelAPI reads
        self.digest = None
        self.api = None
        self.keras_config = None
        self.weights = None

    def get_pipes(self, num=1):
        if self.api is None:
            self.api = ChessModelAPI(self, transport=self.transport)
            self.api.start()
        return [self.api.create_pipe() for _ in range(num)]

    def _install(self, cfg, weights):
        if self.model is not None and self.keras_config == cfg:
            self.model.reload(weights)            # hot reload: handle, stream and pipes stay open
        else:
            if self.model is not None:
                self.model.close()
            self.model = Evaluator(cfg, weights, device=self.device, precision=self.precision,
                                   max_batch=self.max_batch)
        self.keras_config, self.weights = cfg, weights

    def build(self, seed=None):
        """Random-init network from ``config.model`` (model_chess.py:57-94)."""
        mc = self.config.model
        cfg = keras_graph.build_config(stem_k=mc.cnn_first_filter_size, filters=mc.cnn_filter_num,
                                       blocks=mc.res_layer_num, block_k=mc.cnn_filter_size,
                                       value_fc=mc.value_fc_size, input_depth=getattr(mc, "input_depth", 18),
                                       n_labels=self.config.n_labels, l2=mc.l2_reg)
        self._install(cfg, keras_graph.init_weights(cfg, seed))

    @staticmethod
    def fetch_digest(weight_path):
        if os.path.exists(weight_path):
            m = hashlib.sha256()
            with open(weight_path, "rb") as f:
                m.update(f.read())
            return m.hexdigest()
        return None

    def load(self, config_path, weight_path):
        """True iff both files exist and were loaded (model_chess.py:121-153)."""
        if os.path.exists(config_path) and os.path.exists(weight_path):
            logger.debug("loading model from %s", config_path)
            cfg = keras_graph.load_config_file(config_path)
            self._install(cfg, keras_graph.load_weight_file(weight_path, cfg))
            self.digest = self.fetch_digest(weight_path)
            logger.debug("loaded model digest = %s", self.digest)
            return True
        logger.debug("model files does not exist at %s and %s", config_path, weight_path)
        return False

    def save(self, config_path, weight_path):
        """Keras get_config() JSON + weights keyed by Keras names (model_chess.py:155-166)."""
        if self.keras_config is None:
            raise RuntimeError("nothing to save: call build() or load() first")
        with open(config_path, "wt") as f:
            json.dump(self.keras_config, f)
        if self.model is not None:
            self.weights = self.model.weights      # model.model.fit(...) may have trained them (worker/optimize.py:88-93)
        keras_graph.save_weight_file(weight_path, self.weights, self.keras_config)
        self.digest = self.fetch_digest(weight_path)
