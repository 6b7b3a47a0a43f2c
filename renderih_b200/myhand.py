"""The `common/myhand` "graph" model variant -- what the reference's `apps/train.py` and `apps/eval_interhand.py` build by default
(`core/lijun_trainer.py:102-113` -> `common/myhand/lijun_model_graph.load_graph_model`), SURVEY.md 8(f) row 1.

Deltas against `models.model.HandNET_GCN`, all served by the same sm_100a kernels:
  * encoder: torchvision ResNet-50 trunk only, no heat-map / dense-pose heads (common/myhand/encoder_lijun.py:62-104);
  * mid: one 1x1 Conv -> ReLU -> BN per trunk level (2048 / 1024 / 512 / 256 -> 256) + global average pool (encoder_lijun.py:107-146);
  * DualGraph blocks without a Laplacian: LN -> ReLU -> Linear -> LN -> ReLU -> Linear + shortcut (model_attn/DualGraph_lijun.py:28-58);
  * inter-hand attention normalises Lf + Rf and attends with each hand's OWN keys over the other hand's values
    (model_attn/inter_attn_lijun.py:79-91);
  * the decoder returns empty MANO lists and no auxiliary maps (decoder_lijun_graph.py:316-320); it owns MANO layers (used by the
    reference's loss / renderer code, not by forward) whose persistent buffers are part of the 1051-key state_dict.
Same `forward(img) -> (result, paramsDict, handDictList, otherInfo)` contract and the same state_dict keys in the same order.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .assets import default_asset_root, load_mano_dict, load_model_assets
from .config import load_cfg
from .manolayer import ManoLayer
from .model import MLP_GraphBlock, ResNetSimple, _cl, _conv_relu_bn, decoder as _decoder_base


class resnet_mid(nn.Module):
    """common/myhand/encoder_lijun.py:107-146"""

    def __init__(self, model_type='resnet50', in_fmapDim=(2048, 1024, 512, 256), out_fmapDim=(256, 256, 256, 256)):
        super().__init__()
        self.convs = nn.ModuleList([nn.Sequential(_cl(nn.Conv2d(in_fmapDim[i], out_fmapDim[i], kernel_size=1, bias=False)),
                                                  nn.ReLU(inplace=True), nn.BatchNorm2d(out_fmapDim[i])) for i in range(len(out_fmapDim))])
        self.global_feature_dim = 512 * 4
        self.fmaps_dim = list(out_fmapDim)

    def get_info(self):
        return {'global_feature_dim': self.global_feature_dim, 'fmaps_dim': self.fmaps_dim}

    def forward(self, img_fmaps, N):
        x1, H1 = img_fmaps[0]
        gf = ops.global_avgpool(x1, N, H1 * H1)
        fmaps = [(_conv_relu_bn(x, seq[0], seq[2], N, H, H, self.training), H) for (x, H), seq in zip(img_fmaps, self.convs)]
        return gf, fmaps


class MANO(nn.Module):
    """common/utils/mano.py:39-79: holder of a ManoLayer plus the 21-joint regressor (tips 745/317/445/556/673, reordered)."""

    def __init__(self, mano_dict):
        super().__init__()
        self.layer = ManoLayer(mano_dict, center_idx=None)
        self.vertex_num, self.joint_num = 778, 21
        self.face = self.layer.faces
        jr = self.layer.J_regressor.numpy()
        tips = np.zeros((5, jr.shape[1]), np.float32)
        for i, v in enumerate((745, 317, 445, 556, 673)):
            tips[i, v] = 1.0
        jr = np.concatenate((jr, tips))[[0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20], :]
        self.joint_regressor = jr
        self.register_buffer('joint_regressor_torch', torch.from_numpy(jr).float())

    def get_3d_joints(self, vertices):
        return torch.einsum('bik,ji->bjk', [vertices, self.joint_regressor_torch])


class decoder(_decoder_base):
    """common/myhand/decoder_lijun_graph.py:150-320"""

    def __init__(self, cfg, mano_left, mano_right, **kw):
        super().__init__(block_cls=MLP_GraphBlock, attn_variant='lijun', mano_lists=False, **kw)
        self.cfg = cfg
        self.mano = bool(getattr(cfg, 'mano_flag', False))
        if self.mano:
            raise NotImplementedError('mano_flag=True (ParamRegressor tail of decoder_lijun_graph.py:226) is not built yet; see DESIGN.md "next"')
        self.mano_left = MANO(mano_left)
        self.mano_left_layer = self.mano_left.layer
        self.mano_right = MANO(mano_right)
        self.mano_right_layer = self.mano_right.layer
        self.left_face = torch.as_tensor(np.asarray(self.mano_left_layer.faces).astype(np.int64))
        self.right_face = torch.as_tensor(np.asarray(self.mano_right_layer.faces).astype(np.int64))
        # decoder_lijun_graph.py:234-236: the released MANO_LEFT has a mirrored first shape direction
        if torch.sum(torch.abs(self.mano_left_layer.shapedirs[:, 0, :] - self.mano_right_layer.shapedirs[:, 0, :])) < 1:
            self.mano_left_layer.shapedirs[:, 0, :] *= -1


class HandNET_GCN(nn.Module):
    """common/myhand/lijun_model_graph.py:19-34"""

    def __init__(self, encoder, mid_model, decoder, cliff=False):
        super().__init__()
        self.encoder = encoder
        self.mid_model = mid_model
        self.decoder = decoder
        self.cliff = cliff

    def forward(self, img):
        if not (isinstance(img, torch.Tensor) and img.is_cuda):
            raise RuntimeError('renderih_b200.myhand.HandNET_GCN runs only on CUDA (sm_100a) tensors; there is no CPU fallback')
        if img.dtype != torch.float32:
            raise RuntimeError('renderih_b200.myhand.HandNET_GCN expects float32 images')
        ops.seed_state.begin_forward()
        if self.training:
            ops.seed_state.advance(img.device)
        N = img.shape[0]
        img_fmaps = self.encoder(img)
        global_feature, fmaps = self.mid_model(img_fmaps, N)
        return self.decoder(global_feature, fmaps)


def load_graph_model(cfg=None, cliff=False, assets=None, mano_assets=None, asset_root=None):
    """`load_graph_model(cfg)` of common/myhand/lijun_model_graph.py:37-70.  cfg: path | CfgNode-like | None (defaults).
    assets / mano_assets: pre-loaded graph / MANO dictionaries (tests use synthetic ones); default = the reference's misc/ files."""
    if cfg is None or isinstance(cfg, str):
        cfg = load_cfg(cfg)
    et = cfg.MODEL.ENCODER_TYPE
    if 'resnet' not in et:
        raise NotImplementedError('myhand graph variant: only the ResNet encoders are built (common/myhand/encoder_lijun.py:328-337)')
    encoder = ResNetSimple(model_type=et, fmapDim=[128, 128, 128, 128], handNum=2, heatmapDim=21, aux_heads=False)
    mid = resnet_mid(model_type=et, in_fmapDim=[2048, 1024, 512, 256], out_fmapDim=cfg.MODEL.DECONV_DIMS)
    a = assets if assets is not None else load_model_assets(cfg, asset_root)
    if mano_assets is None:
        root = asset_root or default_asset_root()
        mano_assets = {s: load_mano_dict(os.path.join(root, 'mano', 'MANO_%s.pkl' % s.upper())) for s in ('left', 'right')}
    info = mid.get_info()
    dec = decoder(cfg, mano_assets['left'], mano_assets['right'],
                  global_feature_dim=info['global_feature_dim'], f_in_Dim=info['fmaps_dim'], f_out_Dim=cfg.MODEL.IMG_DIMS,
                  gcn_in_dim=cfg.MODEL.GCN_IN_DIM, gcn_out_dim=cfg.MODEL.GCN_OUT_DIM, graph_k=cfg.MODEL.graph_k,
                  graph_layer_num=cfg.MODEL.graph_layer_num, vertex_num=778, dense_coor=a['dense_coor'],
                  left_graph_dict=a['left_graph'], right_graph_dict=a['right_graph'], num_attn_heads=4,
                  upsample_weight=a['upsample'], dropout=cfg.TRAIN.dropout)
    return HandNET_GCN(encoder, mid, dec, cliff)
