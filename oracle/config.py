"""Shape configurations of the MESM inference path (oracle side).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The numbers restate the reference's shipped JSON configs:
  config/QVHighlights/C+SF_C.json, config/charades/C+SF_C.json,
  config/charades/VGG_GloVe.json, config/TACoS/C3D_GloVe.json
(`v_feat_dim` is +2 when `use_tef`, utils/config.py:242-243).
"""
from dataclasses import dataclass, asdict


@dataclass(frozen=True)
class OracleConfig:
    name: str
    dataset_name: str            # selects the SS-MESM video batching branch (model/model.py:185-197)
    v_feat_dim: int              # incl. the 2 tef columns
    t_feat_dim: int
    max_video_l: int
    max_words_l: int
    share_mlp: bool = True       # False -> T2VEncoder_TwoMLP for the enhance encoder (runner.py:190-210)
    hidden_dim: int = 256
    nheads: int = 8
    dim_feedforward: int = 1024
    num_recfw_layers: int = 2
    t2v_layers: int = 2
    enc_layers: int = 2
    dec_layers: int = 2
    num_recss_layers: int = 4
    num_queries: int = 10
    n_input_proj: int = 2
    vocab_size: int = 1111
    rec_fw: bool = True
    rec_ss: bool = True
    aux_loss: bool = True
    recss_tau: float = 0.5
    clip_len: float = 2.0
    max_ts_val: float = 150.0

    def asdict(self):
        return asdict(self)


CONFIGS = {
    "qvhighlights": OracleConfig("qvhighlights", "qvhighlights", 2818, 512, 75, 32, clip_len=2.0, max_ts_val=150.0),
    "charades_csf": OracleConfig("charades_csf", "charades", 2818, 512, 194, 16, clip_len=1.0, max_ts_val=150.0),
    "charades_vgg": OracleConfig("charades_vgg", "charades", 4098, 300, 600, 16, clip_len=0.17, max_ts_val=150.0,
                                 vocab_size=1111),
    "tacos": OracleConfig("tacos", "tacos", 4098, 300, 600, 16, share_mlp=False, clip_len=-1.0, max_ts_val=1000.0),
    # tiny shape for fast CPU tests (not a reference config): same layer counts, small feature dims
    "tiny": OracleConfig("tiny", "charades", 130, 64, 24, 8, clip_len=1.0, max_ts_val=150.0, vocab_size=17),
    "tiny_qvh": OracleConfig("tiny_qvh", "qvhighlights", 130, 64, 24, 8, clip_len=2.0, max_ts_val=150.0, vocab_size=17),
    "tiny_twomlp": OracleConfig("tiny_twomlp", "tacos", 130, 64, 24, 8, share_mlp=False, clip_len=-1.0,
                                max_ts_val=1000.0, vocab_size=17),
}
