"""Bridge to the UNMODIFIED reference (adwardlee/RenderIH) for oracle validation only.

TEST INFRASTRUCTURE -- never imported by the product package `renderih_b200`.

This module imports the reference's own Python modules *in place* from /root/reference
(read-only, present only in the authoring container) or, where that does not exist (the GPU box),
from the byte-for-byte staged copy `oracle/_ref/src` made by `oracle/build_ref.py`,
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

HERE = os.path.dirname(os.path.abspath(__file__))
ASSET_DIR = os.path.join(HERE, '_ref', 'misc')
STAGED_ROOT = os.path.join(HERE, '_ref', 'src')       # byte-for-byte copy made by oracle/build_ref.py (git-ignored; travels to the GPU box)


def _pick_root():
    env = os.environ.get('RIH_REFERENCE_ROOT')
    if env:
        return env
    if os.path.isdir('/root/reference/models'):
        return '/root/reference'
    return STAGED_ROOT


REF_ROOT = _pick_root()


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
                kw.pop('weights', None)         # the product's own containers call resnetXX(weights=None) through the same (patched) symbol
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


# ----------------------------------------------------------------------------- common/myhand variants
MISSING_PACKAGES = ('manopth', 'mmcv', 'pytorch3d', 'imgaug', 'timm', 'tensorboardX', 'matplotlib', 'skimage', 'trimesh', 'fvcore', 'sdf',
                    'tkinter', 'chumpy')


def _install_missing_package_stubs():
    """Third-party packages the reference imports at module level along the constructor path but never calls on the model's forward /
    backward path (SURVEY 8c lists them as absent from this container): any `import pkg.sub` / `from pkg.sub import name` resolves to an
    inert stand-in.  Packages that ARE installed are left alone."""
    import importlib.abc
    import importlib.machinery
    import importlib.util

    class _Dummy:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Dummy()

        def __getattr__(self, name):
            if name.startswith('__'):
                raise AttributeError(name)
            return _Dummy()

    class _StubModule(types.ModuleType):
        __path__ = []

        def __getattr__(self, name):
            if name.startswith('__'):
                raise AttributeError(name)
            return type(name, (_Dummy,), {})

    class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        _rih_stub_finder = True

        def __init__(self, roots):
            self.roots = roots

        def find_spec(self, fullname, path=None, target=None):
            if fullname.split('.')[0] in self.roots:
                return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
            return None

        def create_module(self, spec):
            return _StubModule(spec.name)

        def exec_module(self, module):
            pass

    if any(getattr(f, '_rih_stub_finder', False) for f in sys.meta_path):
        return
    roots = set()
    for name in MISSING_PACKAGES:
        if name in sys.modules:
            continue
        try:
            found = importlib.util.find_spec(name) is not None
        except Exception:
            found = False
        if not found:
            roots.add(name)
    sys.meta_path.append(_Finder(roots))


def build_reference_myhand_model(asset_dir, variant='graph', dropout=0.05, mano_flag=False):
    """`common/myhand/lijun_model_graph.load_graph_model` (variant 'graph', what apps/train.py and apps/eval_interhand.py build by
    default: core/lijun_trainer.py:102-113) on the CPU.  The reference hard-codes `.cuda()` in its decoder constructor
    (decoder_lijun_graph.py:228-232), imports `manopth` and creates output folders at import (main/config.py:132-135); those three
    environment dependencies are shimmed IN MEMORY for the duration of the call (identity `.cuda()`, empty `manopth` module,
    no-op folder creation) -- no reference file is modified or copied, the code that runs is the reference's own.
    MANO pickles are read from `<asset_dir>/mano` through a temporary working directory (main/config.py:123 uses a relative path)."""
    import contextlib
    import torch
    ns = import_reference()

    @contextlib.contextmanager
    def shims():
        saved = (torch.Tensor.cuda, torch.nn.Module.cuda, os.makedirs, os.getcwd())
        tmp = os.path.join(asset_dir, '_cwd')
        os.makedirs(os.path.join(tmp, 'misc'), exist_ok=True)
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        orig_to = torch.Tensor.to

        def is_cuda_dev(d):
            return (isinstance(d, str) and d.startswith('cuda')) or (isinstance(d, torch.device) and d.type == 'cuda')

        def to_cpu(self, *a, **k):        # `.to('cuda')` of common/utils/mano.py:34 (Jr) and friends
            a = tuple('cpu' if is_cuda_dev(x) else x for x in a)
            if is_cuda_dev(k.get('device')):
                k['device'] = 'cpu'
            return orig_to(self, *a, **k)
        torch.Tensor.to = to_cpu
        os.makedirs = lambda *a, **k: None
        _install_missing_package_stubs()
        link = os.path.join(tmp, 'misc', 'mano')
        if not os.path.exists(link):
            os.symlink(os.path.join(asset_dir, 'mano'), link)
        os.chdir(tmp)
        try:
            yield
        finally:
            torch.Tensor.cuda, torch.nn.Module.cuda, os.makedirs = saved[:3]
            torch.Tensor.to = orig_to
            os.chdir(saved[3])

    with shims():
        import main.config as main_config
        main_config.cfg.mano_flag = mano_flag
        for k, v in (('render', False), ('normal', True), ('edge', True), ('vert2d', True), ('dice', False), ('sdf', False)):
            if not hasattr(main_config.cfg, k):
                setattr(main_config.cfg, k, v)
        # common/utils/mano_to_vertex.py pulls in the dead common/nets harness, which instantiates a manopth layer at import
        # (common/nets/mano_head.py:8); it is only needed by the unused CLIFF bbox decoder (bbox_decoder.py:22) -> inert stand-in
        if 'common.utils.mano_to_vertex' not in sys.modules:
            m = types.ModuleType('common.utils.mano_to_vertex')
            m.mano_convert = None
            sys.modules['common.utils.mano_to_vertex'] = m
        paths = {'left': os.path.join(asset_dir, 'graph_left.pkl'), 'right': os.path.join(asset_dir, 'graph_right.pkl')}
        if variant == 'graph':
            import common.myhand.lijun_model_graph as mod
            import common.myhand.decoder_lijun_graph as dec
            loader = 'load_graph_model'
        elif variant == 'newgraph':      # lijun_model_newgraph.load_new_model + decoder_lijun_mano (MANO tail)
            import common.myhand.lijun_model_newgraph as mod
            import common.myhand.decoder_lijun_mano as dec
            loader = 'load_new_model'
            main_config.cfg.mano_flag = True
            if not hasattr(main_config.cfg, 'reverse'):
                main_config.cfg.reverse = False
        else:
            raise ValueError(variant)
        dec.get_graph_dict_path = lambda: paths
        dec.get_dense_color_path = lambda: os.path.join(asset_dir, 'v_color.pkl')
        dec.get_upsample_path = lambda: os.path.join(asset_dir, 'upsample.pkl')
        cfg = mod.load_cfg()
        cfg.TRAIN.dropout = dropout
        cfg.MODEL_PARAM.MODEL_PRETRAIN_PATH = '__none__'
        model = getattr(mod, loader)(cfg)
    model._rih_cpu_shims = shims          # the 'newgraph' forward calls .cuda() too (decoder_lijun_mano.py:51): run it under `with model._rih_cpu_shims():`
    return model, cfg
