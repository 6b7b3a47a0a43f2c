"""renderih_b200 -- B200-native (sm_100a) hot path of adwardlee/RenderIH: HandNET_GCN forward/backward + ManoLayer.

Public API mirrors the reference modules it replaces:
    load_model(cfg) / HandNET_GCN      <-  models/model.py:18-60
    ManoLayer, rodrigues_batch          <-  models/manolayer.py:32-48,100-322
    load_graph_model(cfg)               <-  common/myhand/lijun_model_graph.py:37-70 (the trainers' default model)
    load_new_model(cfg)                 <-  common/myhand/lijun_model_newgraph.py:34-71 (graph model + MANO tail)
    preprocess_u8(frames, flip)         <-  core/loader.py:151-152,178-181 (host image ops of the loader, on the GPU)
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
    if name in ('load_graph_model', 'load_new_model'):
        from . import myhand
        return getattr(myhand, name)
    if name == 'preprocess_u8':
        from . import input as _input
        return _input.preprocess_u8
    if name in ('load_cfg', 'get_cfg_defaults'):
        from . import config
        return getattr(config, name)
    raise AttributeError(name)
