"""Global configuration object, same keys and defaults as the reference
(``miscc/config.py:9-66``) and the same YAML merge rules (key must exist, types must match).
Works with or without the ``easydict`` package."""
import numpy as np

try:
    from easydict import EasyDict as edict
except ImportError:
    class edict(dict):
        """minimal attribute-access dict (stand-in for easydict.EasyDict)"""

        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                self[k] = v

        def __setitem__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, edict):
                v = edict(v)
            super().__setitem__(k, v)

        __setattr__ = __setitem__

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)


__C = edict()
cfg = __C

_DEFAULTS = dict(
    DATASET_NAME="birds", EMBEDDING_TYPE="cnn-rnn", CONFIG_NAME="", GPU_ID="0", CUDA=True, WORKERS=6,
    VIDEO_LEN=5, NET_G="", NET_D="", STAGE1_G="", DATA_DIR="", VIS_COUNT=64,
    USE_SEQ_CONSISTENCY=False, CONSISTENCY_RATIO=1.0, SEGMENT_LEARNING=True, SEGMENT_RATIO=1.0,
    IMAGE_RATIO=5.0, RECONSTRUCT_LOSS=1.0, EVALUATE_FID_SCORE=False, CASCADE_MODEL=True,
    Z_DIM=100, IMSIZE=64, SESIZE=64, STAGE=1, LABEL_NUM=9,
    TRAIN=dict(FLAG=True, IM_BATCH_SIZE=64, ST_BATCH_SIZE=64, MAX_EPOCH=600, SNAPSHOT_INTERVAL=50,
               PRETRAINED_MODEL="", PRETRAINED_EPOCH=600, LR_DECAY_EPOCH=600, DISCRIMINATOR_LR=2e-4,
               GENERATOR_LR=2e-4, SEGMENT_NAME="img_segment", COEFF=dict(KL=2.0)),
    GAN=dict(CONDITION_DIM=124, Z_DIM=100, DF_DIM=124, GF_DIM=256, GF_SEG_DIM=1024, R_NUM=4),
    TEXT=dict(DIMENSION=356),
)
for _k, _v in _DEFAULTS.items():
    __C[_k] = _v


def _merge_a_into_b(a, b):
    if not isinstance(a, dict):
        return
    for k, v in a.items():
        if k not in b:
            raise KeyError("{} is not a valid config key".format(k))
        old = b[k]
        if type(old) is not type(v) and not (isinstance(old, dict) and isinstance(v, dict)):
            if isinstance(old, np.ndarray):
                v = np.array(v, dtype=old.dtype)
            elif isinstance(old, float) and isinstance(v, int):
                v = float(v)
            else:
                raise ValueError("Type mismatch ({} vs. {}) for config key: {}".format(type(old), type(v), k))
        if isinstance(v, dict):
            _merge_a_into_b(v, b[k])
        else:
            b[k] = v


def cfg_from_file(filename):
    """Load a YAML config file and merge it into the defaults."""
    import yaml
    with open(filename, "r") as f:
        yaml_cfg = yaml.safe_load(f)
    _merge_a_into_b(yaml_cfg, __C)
