"""The hot-path constants of the reference's configuration, with the same attribute names
(/root/reference/src/chess_zero/configs/normal.py:28-43,60-68; configs/mini.py:31,34 differ only in
max_processes and simulation_num_per_move).  Only what the evaluation path reads is mirrored; the
reference's own ``Config`` object can be passed instead wherever a config is accepted."""
from . import labels as _labels


class ModelConfig:
    cnn_filter_num = 256
    cnn_first_filter_size = 5
    cnn_filter_size = 3
    res_layer_num = 7
    l2_reg = 1e-4
    value_fc_size = 256
    distributed = False
    input_depth = 18


class PlayConfig:
    max_processes = 3
    search_threads = 16
    simulation_num_per_move = 800
    c_puct = 1.5
    noise_eps = 0.25
    dirichlet_alpha = 0.3
    virtual_loss = 3


class Config:
    labels = _labels.LABELS
    n_labels = _labels.N_LABELS
    flipped_labels = _labels.FLIPPED_LABELS
    unflipped_index = _labels.UNFLIPPED_INDEX

    def __init__(self, config_type="normal"):
        self.model = ModelConfig()
        self.play = PlayConfig()
        if config_type == "mini":
            self.play.max_processes = 1
            self.play.simulation_num_per_move = 100
        elif config_type not in ("normal", "distributed"):
            raise RuntimeError(f"unknown config_type: {config_type}")

    @staticmethod
    def flip_policy(pol):
        return _labels.flip_policy(pol)
