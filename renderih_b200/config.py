"""Minimal attribute-access config mirroring the reference's yacs CfgNode usage (utils/config.py:7-21).

Defaults restate the hot-path-relevant keys of the reference's utils/defaults.yaml:12-21,39-49.
Any object with the same attribute layout (e.g. a real yacs CfgNode) is accepted by load_model().
"""
import copy

import yaml

DEFAULTS = {
    'SEED': 88,
    'MISC': {
        'MANO_PATH': 'misc/mano',
        'GRAPH_LEFT_DICT_PATH': 'misc/graph_left.pkl',
        'GRAPH_RIGHT_DICT_PATH': 'misc/graph_right.pkl',
        'DENSE_COLOR': 'misc/v_color.pkl',
        'UPSAMPLE_PATH': 'misc/upsample.pkl',
    },
    'MODEL': {
        'ENCODER_TYPE': 'resnet50',
        'DECONV_DIMS': [256, 256, 256, 256],
        'IMG_DIMS': [256, 128, 64],
        'GCN_IN_DIM': [512, 256, 128],
        'GCN_OUT_DIM': [256, 128, 64],
        'ENCODER_PRETRAIN_PATH': 'none',
        'freeze_upsample': True,
        'graph_k': 2,
        'graph_layer_num': 4,
    },
    'MODEL_PARAM': {'MODEL_PRETRAIN_PATH': 'none'},
    'TRAIN': {'BATCH_SIZE': 64, 'LR': 3.0e-4, 'dropout': 0.05, 'weight_decay': 1.0e-2, 'OPTIM': 'adam',
              'EPOCHS': 200, 'current_epoch': 0, 'lr_decay_step': 80, 'lr_decay_gamma': 0.1, 'warm_up': 3},     # utils/defaults.yaml:36-44
    'LOSS_WEIGHT': {
        'DATA': {'LABEL_3D': 100, 'LABEL_2D': 50, 'MANO_POSE': 0.5, 'MANO_SHAPE': 0.01, 'MANO_REL': 1},     # utils/defaults.yaml:54-61
        'GRAPH': {'NORM': {'EDGE': 2000, 'NORMAL': 10, 'NORM_EPOCH': 50}},
        'NORM': {'UPSAMPLE': 1.0},
    },
}


class CfgNode(dict):
    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) else copy.deepcopy(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def merge(self, d):
        for k, v in d.items():
            if isinstance(v, dict) and isinstance(self.get(k), CfgNode):
                self[k].merge(v)
            else:
                self[k] = CfgNode(v) if isinstance(v, dict) else v

    def merge_from_file(self, path):
        with open(path) as f:
            self.merge(yaml.safe_load(f) or {})

    def clone(self):
        return CfgNode(self)


def get_cfg_defaults():
    return CfgNode(DEFAULTS)


def load_cfg(path=None):
    cfg = get_cfg_defaults()
    if path is not None:
        cfg.merge_from_file(path)
    return cfg
