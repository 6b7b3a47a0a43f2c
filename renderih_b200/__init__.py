"""renderih_b200 -- B200-native (sm_100a) hot path of adwardlee/RenderIH: HandNET_GCN forward/backward + ManoLayer.

Public API mirrors the reference modules it replaces:
    load_model(cfg) / HandNET_GCN      <-  models/model.py:18-60
    ManoLayer, rodrigues_batch          <-  models/manolayer.py:32-48,100-322
    load_graph_model(cfg)               <-  common/myhand/lijun_model_graph.py:37-70 (the trainers' default model)
    load_new_model(cfg)                 <-  common/myhand/lijun_model_newgraph.py:34-71 (graph model + MANO tail)
    preprocess_u8(frames, flip)         <-  core/loader.py:151-152,178-181 (host image ops of the loader, on the GPU)
    augment_u8 / get_affine_mat / prepare_labels  <-  core/loader.py:96-211, utils/manoutils.py:183-261 (training-time augmentation:
                                            cv2.warpAffine, brightness noise, flip, normalisation in one kernel; label transforms)
    calc_loss_GCN / mano_loss_GCN       <-  core/Loss.py:201-277, core/Loss_mano.py:245-335 (fused mesh-term kernels, loss.py)
    Jr, batch_metrics, EvalMetrics      <-  apps/eval_interhand.py:28-170,300-552, utils/eval_metrics.py:36-50 (metrics.py)
    TrainStep, lr_at_epoch              <-  core/lijun_trainer.py:131-159,262-313, utils/lr_sc.py:159-175 (train.py)
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
    if name in ('preprocess_u8', 'augment_u8', 'get_affine_mat', 'augment_labels', 'prepare_labels'):
        from . import input as _input
        return getattr(_input, name)
    if name in ('GraphLoss', 'ManoLoss', 'calc_loss_GCN', 'mano_loss_GCN'):
        from . import loss
        return getattr(loss, name)
    if name in ('Jr', 'batch_metrics', 'EvalMetrics'):
        from . import metrics
        return getattr(metrics, name)
    if name in ('TrainStep', 'FlatParams', 'lr_at_epoch'):
        from . import train
        return getattr(train, name)
    if name in ('load_cfg', 'get_cfg_defaults'):
        from . import config
        return getattr(config, name)
    raise AttributeError(name)
