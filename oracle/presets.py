"""Workload presets (BASELINE.json ``configs``).

Every preset is a flat dict of the ``cfg`` fields the model reads at construction time
(reference ``miscc/config.py:9-66``, ``cfg/final.yml``) plus the two batch sizes.
"""

_COMMON = dict(
    CUDA=False, USE_SEQ_CONSISTENCY=False, SEGMENT_LEARNING=True, CASCADE_MODEL=False,
    SEGMENT_RATIO=1.0, IMAGE_RATIO=5.0, KL=1.0, RECONSTRUCT_LOSS=1.0, CONSISTENCY_RATIO=1.0,
    DISCRIMINATOR_LR=4e-4, GENERATOR_LR=1e-4,
)

PRESETS = {
    # cfg/final.yml -- BASELINE.json configs[1]
    "pororo": dict(_COMMON, VIDEO_LEN=5, TEXT_DIM=356, LABEL_NUM=9, ST_BATCH=18, IM_BATCH=90,
                   CONDITION_DIM=124, Z_DIM=100, DF_DIM=124, GF_DIM=256, GF_SEG_DIM=1024),
    # BASELINE.json configs[0]: CLEVR-shaped stories, B=4, 4 frames (datasets/clevr.py:24,40-41)
    "clevr": dict(_COMMON, VIDEO_LEN=4, TEXT_DIM=72, LABEL_NUM=15, ST_BATCH=4, IM_BATCH=16,
                  CONDITION_DIM=124, Z_DIM=100, DF_DIM=124, GF_DIM=256, GF_SEG_DIM=1024),
    # BASELINE.json configs[4]: large-batch stress
    "stress": dict(_COMMON, VIDEO_LEN=5, TEXT_DIM=356, LABEL_NUM=9, ST_BATCH=512, IM_BATCH=2560,
                   CONDITION_DIM=124, Z_DIM=100, DF_DIM=124, GF_DIM=256, GF_SEG_DIM=1024),
    # reduced-width model used for the committed golden fixtures (full gradients fit in <2 MB)
    "tiny": dict(_COMMON, VIDEO_LEN=3, TEXT_DIM=20, LABEL_NUM=3, ST_BATCH=2, IM_BATCH=6,
                 CONDITION_DIM=12, Z_DIM=10, DF_DIM=8, GF_DIM=8, GF_SEG_DIM=32),
    # same widths as the real model in the channel dims that matter for tiling, small batch
    "small": dict(_COMMON, VIDEO_LEN=3, TEXT_DIM=40, LABEL_NUM=5, ST_BATCH=4, IM_BATCH=8,
                  CONDITION_DIM=28, Z_DIM=16, DF_DIM=31, GF_DIM=32, GF_SEG_DIM=128),
}
# the cascade generator (cascade_model.py, cfg.CASCADE_MODEL)
for _n in ("tiny", "small", "clevr", "pororo"):
    PRESETS[_n + "_cascade"] = dict(PRESETS[_n], CASCADE_MODEL=True)
# the order-consistency critic inside the story discriminator (cfg.USE_SEQ_CONSISTENCY)
for _n in ("tiny", "small", "clevr", "pororo"):
    PRESETS[_n + "_seq"] = dict(PRESETS[_n], USE_SEQ_CONSISTENCY=True)


def get(name, **overrides):
    p = dict(PRESETS[name])
    p.update(overrides)
    p["name"] = name
    return p


def apply_to_cfg(cfg, p):
    """Write preset ``p`` into a reference-style ``cfg`` EasyDict (miscc/config.py:9-66)."""
    cfg.CUDA = p["CUDA"]
    cfg.VIDEO_LEN = p["VIDEO_LEN"]
    cfg.LABEL_NUM = p["LABEL_NUM"]
    cfg.USE_SEQ_CONSISTENCY = p["USE_SEQ_CONSISTENCY"]
    cfg.SEGMENT_LEARNING = p["SEGMENT_LEARNING"]
    cfg.CASCADE_MODEL = p["CASCADE_MODEL"]
    cfg.SEGMENT_RATIO = p["SEGMENT_RATIO"]
    cfg.IMAGE_RATIO = p["IMAGE_RATIO"]
    cfg.RECONSTRUCT_LOSS = p.get("RECONSTRUCT_LOSS", 1.0)
    cfg.CONSISTENCY_RATIO = p.get("CONSISTENCY_RATIO", 1.0)
    cfg.Z_DIM = p["Z_DIM"]
    cfg.TRAIN.IM_BATCH_SIZE = p["IM_BATCH"]
    cfg.TRAIN.ST_BATCH_SIZE = p["ST_BATCH"]
    cfg.TRAIN.DISCRIMINATOR_LR = p["DISCRIMINATOR_LR"]
    cfg.TRAIN.GENERATOR_LR = p["GENERATOR_LR"]
    cfg.TRAIN.COEFF.KL = p["KL"]
    cfg.GAN.CONDITION_DIM = p["CONDITION_DIM"]
    cfg.GAN.Z_DIM = p["Z_DIM"]
    cfg.GAN.DF_DIM = p["DF_DIM"]
    cfg.GAN.GF_DIM = p["GF_DIM"]
    cfg.GAN.GF_SEG_DIM = p["GF_SEG_DIM"]
    cfg.TEXT.DIMENSION = p["TEXT_DIM"]
    return cfg
