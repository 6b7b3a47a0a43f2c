"""Bridge to the UNMODIFIED reference (adwardlee/RenderIH) for oracle validation only.

TEST INFRASTRUCTURE -- never imported by the product package `renderih_b200`.

This module imports the reference's own Python modules *in place* from /root/reference
(read-only, present only in the authoring container; it does not exist on the GPU box),
after installing two tiny import shims for packages the container lacks:

  * `yacs.config.CfgNode`   (needed by  utils/config.py:1)
  * `chumpy`                (needed to unpickle MANO `shapedirs`, models/manolayer.py:141-144)

and extracts the reference's asset bundle `misc.tar` (actually a ZIP) into
`oracle/_ref/misc/` (git-ignored, travels to the GPU box with gpurun).  No reference
source file is copied into this repository.
"""
import os
import sys
import types
import zipfile
import pickle

import numpy as np

REF_ROOT = os.environ.get('RIH_REFERENCE_ROOT', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
ASSET_DIR = os.path.join(HERE, '_ref', 'misc')


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, 'models'))


def extract_assets(force=False):
    """Unpack misc.tar (a ZIP) to oracle/_ref/misc/."""
    marker = os.path.join(ASSET_DIR, 'graph_left.pkl')
    if os.path.exists(marker) and not force:
        return ASSET_DIR
    src = os.path.join(REF_ROOT, 'misc.tar')
    os.makedirs(os.path.dirname(ASSET_DIR), exist_ok=True)
    with zipfile.ZipFile(src) as z:
        for info in z.infolist():
            if info.filename.startswith('misc/') and not info.is_dir():
                z.extract(info, os.path.dirname(ASSET_DIR))
    return ASSET_DIR


# ----------------------------------------------------------------------------- shims
def _install_yacs_shim():
    if 'yacs.config' in sys.modules:
        return
    import yaml

    class CfgNode(dict):
        def __init__(self, init_dict=None, key_list=None, new_allowed=False):
            super().__init__()
            for k, v in (init_dict or {}).items():
                self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

        def merge_from_file(self, path):
            with open(path) as f:
                self._merge(yaml.safe_load(f))

        def merge_from_other_cfg(self, other):
            self._merge(other)

        def _merge(self, d):
            for k, v in d.items():
                if isinstance(v, dict):
                    if k not in self or not isinstance(self[k], CfgNode):
                        self[k] = CfgNode()
                    self[k]._merge(v)
                else:
                    self[k] = v

        def set_new_allowed(self, flag):
            pass

        def clone(self):
            import copy
            return copy.deepcopy(self)

        def freeze(self):
            pass

        def defrost(self):
            pass

        def dump(self, **kw):
            def plain(n):
                return {k: plain(v) if isinstance(v, dict) else v for k, v in n.items()}
            return yaml.safe_dump(plain(self))

    yacs = types.ModuleType('yacs')
    cfgmod = types.ModuleType('yacs.config')
    cfgmod.CfgNode = CfgNode
    yacs.config = cfgmod
    sys.modules['yacs'] = yacs
    sys.modules['yacs.config'] = cfgmod


def _install_chumpy_shim():
    if 'chumpy' in sys.modules:
        return

    class Ch(object):
        def __setstate__(self, state):
            self.__dict__.update(state)

        @property
        def r(self):
            return np.asarray(self.x)

    class Select(Ch):
        @property
        def r(self):
            a = self.a.r if hasattr(self.a, 'r') else np.asarray(self.a)
            return a.ravel()[np.asarray(self.idxs)].reshape(self.preferred_shape)

    chumpy = types.ModuleType('chumpy')
    ch = types.ModuleType('chumpy.ch')
    reordering = types.ModuleType('chumpy.reordering')
    ch.Ch = Ch
    reordering.Select = Select
    chumpy.ch = ch
    chumpy.reordering = reordering
    chumpy.Ch = Ch
    sys.modules['chumpy'] = chumpy
    sys.modules['chumpy.ch'] = ch
    sys.modules['chumpy.reordering'] = reordering


_imported = {}


def import_reference():
    """Return a namespace with the reference modules of the hot path (imported in place)."""
    if _imported:
        return _imported['ns']
    assert reference_available(), 'reference tree not present (only exists in the authoring container)'
    _install_yacs_shim()
    _install_chumpy_shim()
    import torchvision.models as tvm
    # load_encoder asks for pretrained=True (models/encoder.py:358); no network.
    for name in ['resnet18', 'resnet34', 'resnet50', 'resnet101', 'resnet152']:
        orig = getattr(tvm, name)
        if getattr(orig, '_rih_patched', False):
            continue

        def make(orig):
            def f(pretrained=False, **kw):
                return orig(weights=None, **kw)
            f._rih_patched = True
            return f
        setattr(tvm, name, make(orig))
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    warnings.filterwarnings('ignore')
    import models.model as ref_model
    import models.decoder as ref_decoder
    import models.encoder as ref_encoder
    import models.manolayer as ref_mano
    import core.Loss as ref_loss
    import utils.config as ref_config
    ns = types.SimpleNamespace(model=ref_model, decoder=ref_decoder, encoder=ref_encoder,
                               mano=ref_mano, loss=ref_loss, config=ref_config)
    _imported['ns'] = ns
    return ns


def build_reference_model(asset_dir=None, encoder_type='resnet50', dropout=0.05):
    """load_model(cfg) of models/model.py:40-60 with asset paths redirected to `asset_dir`."""
    ns = import_reference()
    asset_dir = asset_dir or extract_assets()
    dec = ns.decoder
    dec.get_graph_dict_path = lambda: {'left': os.path.join(asset_dir, 'graph_left.pkl'),
                                       'right': os.path.join(asset_dir, 'graph_right.pkl')}
    dec.get_dense_color_path = lambda: os.path.join(asset_dir, 'v_color.pkl')
    dec.get_upsample_path = lambda: os.path.join(asset_dir, 'upsample.pkl')
    ns.loss.get_upsample_path = dec.get_upsample_path
    cfg = ns.config.load_cfg()
    cfg.MODEL.ENCODER_TYPE = encoder_type
    cfg.TRAIN.dropout = dropout
    return ns.model.load_model(cfg), cfg


def build_reference_mano(side='right', asset_dir=None, **kw):
    ns = import_reference()
    asset_dir = asset_dir or extract_assets()
    path = os.path.join(asset_dir, 'mano', 'MANO_%s.pkl' % side.upper())
    return ns.mano.ManoLayer(path, **kw)
