"""renderih_b200 -- B200-native (sm_100a) hot path of adwardlee/RenderIH: HandNET_GCN forward/backward + ManoLayer.

Public API mirrors the reference modules it replaces:
    load_model(cfg) / HandNET_GCN      <-  models/model.py:18-60
    ManoLayer, rodrigues_batch          <-  models/manolayer.py:32-48,100-322
"""
from . import _build


def build(force=False):
    """Compile the CUDA library for sm_100a (in-tree)."""
    return _build.build(force=force)


def __getattr__(name):
    if name in ('load_model', 'HandNET_GCN', 'load_encoder', 'load_decoder'):
        from . import model
        return getattr(model, name)
    if name in ('ManoLayer', 'rodrigues_batch'):
        from . import manolayer
        return getattr(manolayer, name)
    if name in ('load_cfg', 'get_cfg_defaults'):
        from . import config
        return getattr(config, name)
    raise AttributeError(name)
