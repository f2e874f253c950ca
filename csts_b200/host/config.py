"""Configuration node and defaults for the keys the CSTS hot path reads.

The reference builds its cfg from fvcore/yacs (``slowfast/config/defaults.py:12-942`` plus
``custom_config.py:8-25``).  Neither package exists in this image, and only ~40 of the ~400 default
keys reach the hot path, so this module provides a small attribute-dict node with the same
surface (``clone``, ``merge_from_file``, ``merge_from_list``, ``dump``) and the defaults of exactly
those keys.  Unknown sections/keys found in a YAML (data-loader, tensorboard, ... settings that the
reference's other subsystems consume) are kept verbatim, so the reference's
``configs/{Ego4D,Aria}/*.yaml`` load unchanged.
"""
import ast
import copy

import yaml


class CfgNode(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    @staticmethod
    def _decode(v):
        # YAML has no tuples: "(3, 7, 7)" arrives as a string (yacs literal_evals it as well)
        if isinstance(v, str):
            try:
                v = ast.literal_eval(v)
            except (ValueError, SyntaxError):
                return v
        if isinstance(v, tuple):
            v = list(v)
        return v

    def _merge(self, d):
        for k, v in d.items():
            if isinstance(v, dict):
                node = self.get(k)
                if not isinstance(node, CfgNode):
                    node = self[k] = CfgNode()
                node._merge(v)
            else:
                self[k] = self._decode(v)

    def merge_from_file(self, path):
        with open(path) as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, opts):
        if len(opts) % 2:
            raise ValueError("override list must be KEY VALUE pairs")
        for key, val in zip(opts[0::2], opts[1::2]):
            node = self
            *parents, leaf = key.split(".")
            for p in parents:
                if p not in node:
                    raise KeyError(f"Non-existent config key: {key}")
                node = node[p]
            if leaf not in node:
                raise KeyError(f"Non-existent config key: {key}")
            node[leaf] = self._decode(val)

    def dump(self, **kw):
        def plain(n):
            return {k: plain(v) for k, v in n.items()} if isinstance(n, dict) else n
        return yaml.safe_dump(plain(self), **kw)


def get_cfg():
    """Defaults of the keys used by build_model / CSTS / the training step.
    Values: slowfast/config/defaults.py (line numbers beside each group) and custom_config.py."""
    c = CfgNode()
    c.TRAIN = CfgNode(dict(ENABLE=True, DATASET="kinetics", BATCH_SIZE=64, MIXED_PRECISION=False,       # :41-79
                           EVAL_PERIOD=10, CHECKPOINT_PERIOD=10, AUTO_RESUME=True, CHECKPOINT_FILE_PATH="",
                           CHECKPOINT_TYPE="pytorch", CHECKPOINT_INFLATE=False, CHECKPOINT_EPOCH_RESET=False,
                           AUDIO_CHECKPOINT_FILE_PATH=""))
    c.TEST = CfgNode(dict(ENABLE=True, DATASET="kinetics", BATCH_SIZE=8, NUM_ENSEMBLE_VIEWS=10, NUM_SPATIAL_CROPS=3))
    c.DATA = CfgNode(dict(PATH_PREFIX="", NUM_FRAMES=8, SAMPLING_RATE=8, TRAIN_JITTER_SCALES=[256, 320],   # :420-470
                          TRAIN_CROP_SIZE=224, TEST_CROP_SIZE=256, INPUT_CHANNEL_NUM=[3, 3], TARGET_FPS=30,
                          USE_OFFSET_SAMPLING=False, MEAN=[0.45, 0.45, 0.45], STD=[0.225, 0.225, 0.225],
                          GAUSSIAN_KERNEL=19))
    c.MVIT = CfgNode(dict(MODE="conv", POOL_FIRST=False, CLS_EMBED_ON=True, AUDIO_BRANCH_ON=False,           # :303-383
                          PATCH_KERNEL=[3, 7, 7], PATCH_STRIDE=[2, 4, 4], PATCH_PADDING=[2, 4, 4], PATCH_2D=False,
                          EMBED_DIM=96, NUM_HEADS=1, MLP_RATIO=4.0, QKV_BIAS=True, DROPPATH_RATE=0.1, DEPTH=16,
                          NORM="layernorm", DIM_MUL=[], HEAD_MUL=[], POOL_KV_STRIDE=None,
                          POOL_KV_STRIDE_ADAPTIVE=None, POOL_Q_STRIDE=[], POOL_KVQ_KERNEL=None,
                          ZERO_DECAY_POS_CLS=True, NORM_STEM=False, SEP_POS_EMBED=False, DROPOUT_RATE=0.0,
                          SPATIAL_AUDIO_ATTN=False))
    c.MODEL = CfgNode(dict(ARCH="slowfast", MODEL_NAME="SlowFast", NUM_CLASSES=400, LOSS_FUNC="cross_entropy",  # :262-297
                           DROPOUT_RATE=0.5, ACT_CHECKPOINT=False, LOSS_ALPHA=1.0))
    c.SOLVER = CfgNode(dict(BASE_LR=0.1, LR_POLICY="cosine", COSINE_END_LR=0.0, MAX_EPOCH=300, MOMENTUM=0.9,   # :499-560
                            WEIGHT_DECAY=1e-4, WARMUP_EPOCHS=0.0, WARMUP_START_LR=0.01, OPTIMIZING_METHOD="sgd",
                            BASE_LR_SCALE_NUM_SHARDS=False, COSINE_AFTER_WARMUP=False, ZERO_WD_1D_PARAM=False,
                            CLIP_GRAD_VAL=None, CLIP_GRAD_L2NORM=None, DAMPENING=0.0, NESTEROV=True))
    c.BN = CfgNode(dict(USE_PRECISE_STATS=False, NUM_BATCHES_PRECISE=200, WEIGHT_DECAY=0.0))
    c.DATA_LOADER = CfgNode(dict(NUM_WORKERS=8, PIN_MEMORY=True, RETURN_TARGET_FRAME=False))
    c.TENSORBOARD = CfgNode(dict(ENABLE=False))
    c.NUM_GPUS = 1          # :566
    c.NUM_SHARDS = 1
    c.SHARD_ID = 0
    c.OUTPUT_DIR = "./tmp"
    c.RNG_SEED = 1
    c.LOG_PERIOD = 10
    c.LOG_MODEL_INFO = True
    c.DIST_BACKEND = "nccl"  # :594
    return c


def assert_and_infer_cfg(cfg):
    """The checks of slowfast/config/defaults.py:945-970 that concern this path."""
    if cfg.NUM_GPUS:
        assert cfg.TRAIN.BATCH_SIZE % cfg.NUM_GPUS == 0
        assert cfg.TEST.BATCH_SIZE % cfg.NUM_GPUS == 0
    assert cfg.NUM_SHARDS > 0 and cfg.SHARD_ID < cfg.NUM_SHARDS
    if cfg.SOLVER.BASE_LR_SCALE_NUM_SHARDS:
        cfg.SOLVER.BASE_LR *= cfg.NUM_SHARDS
    return cfg
