"""Asset loading (graph Laplacians / permutations / upsample matrix / dense colours / MANO tensors).

The reference reads these pickles from `misc/` (models/decoder.py:177-193, models/manolayer.py:101-152,
paths from utils/manoutils.py:68-103).  They are NOT shipped in this repository (MANO is licence-restricted):
point `asset_root` / $RIH_ASSET_ROOT at a directory containing `misc/...`, e.g. the reference checkout after
unpacking its `misc.tar`.  `synthetic_assets()` builds shape-compatible random stand-ins for tests and benchmarks.
"""
import os
import pickle
import sys
import types

import numpy as np
import scipy.sparse as sp


def default_asset_root():
    env = os.environ.get('RIH_ASSET_ROOT')
    if env:
        return env
    here = os.path.dirname(os.path.abspath(__file__))
    cand = os.path.join(here, '..', 'oracle', '_ref')
    if os.path.exists(os.path.join(cand, 'misc', 'graph_left.pkl')):
        return os.path.abspath(cand)
    return os.getcwd()


def _install_chumpy_shim():
    """MANO pickles embed a chumpy object for `shapedirs` (manolayer.py:141-144); restore it without chumpy."""
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

    chumpy = types.ModuleType('chumpy'); ch = types.ModuleType('chumpy.ch'); ro = types.ModuleType('chumpy.reordering')
    ch.Ch = Ch; ro.Select = Select; chumpy.ch = ch; chumpy.reordering = ro; chumpy.Ch = Ch
    sys.modules['chumpy'] = chumpy; sys.modules['chumpy.ch'] = ch; sys.modules['chumpy.reordering'] = ro


def load_pickle(path, latin1=False):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        with open(path, 'rb') as f:
            return pickle.load(f, encoding='latin1') if latin1 else pickle.load(f)


def load_model_assets(cfg, asset_root=None):
    """-> dict(left_graph, right_graph, dense_coor, upsample) exactly as load_decoder reads them."""
    root = asset_root or default_asset_root()
    out = {}
    out['left_graph'] = load_pickle(os.path.join(root, cfg.MISC.GRAPH_LEFT_DICT_PATH))
    out['right_graph'] = load_pickle(os.path.join(root, cfg.MISC.GRAPH_RIGHT_DICT_PATH))
    out['dense_coor'] = np.asarray(load_pickle(os.path.join(root, cfg.MISC.DENSE_COLOR)), dtype=np.float32)
    out['upsample'] = np.asarray(load_pickle(os.path.join(root, cfg.MISC.UPSAMPLE_PATH)), dtype=np.float32)
    return out


def load_mano_dict(mano_path):
    _install_chumpy_shim()
    d = load_pickle(mano_path, latin1=True)
    sd = d['shapedirs']
    if not isinstance(sd, np.ndarray):
        sd = np.array(sd.r)
    out = dict(d)
    out['shapedirs'] = np.asarray(sd, dtype=np.float32)
    return out


# ----------------------------------------------------------------------------- synthetic stand-ins
def _ring_laplacian(n, rng, extra=3):
    """Symmetric rescaled-Laplacian-like sparse matrix with <= ~11 nnz/row (same structure class as misc/graph_*.pkl)."""
    rows, cols = [], []
    for i in range(n):
        for d in (1, 2):
            rows += [i, (i + d) % n]; cols += [(i + d) % n, i]
        for _ in range(extra // 2):
            j = int(rng.randint(0, n))
            if j != i:
                rows += [i, j]; cols += [j, i]
    A = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n))
    A.data[:] = 1.0
    A = A.tocsr(); A.sum_duplicates(); A.data[:] = 1.0
    d = np.asarray(A.sum(1)).ravel()
    dinv = 1.0 / np.sqrt(np.maximum(d, 1))
    L = sp.eye(n) - sp.diags(dinv) @ A @ sp.diags(dinv)
    return sp.csr_matrix(L - sp.eye(n))  # rescaled: 2L/lmax - I with lmax=2


def synthetic_assets(seed=0):
    """Random assets with the real shapes: graphs 1008/504/252/126/63, perm 1008, upsample 778x252, colours 778x3."""
    rng = np.random.RandomState(seed)
    out = {}
    for side in ('left', 'right'):
        Ls = [_ring_laplacian(n, rng).astype(np.float64) for n in (1008, 504, 252, 126, 63)]
        perm = np.concatenate([rng.permutation(778), rng.randint(0, 778, size=1008 - 778)])
        rng.shuffle(perm)
        perm_rev = np.zeros(1008, dtype=np.int64)
        first = {}
        for pos, v in enumerate(perm):
            first.setdefault(int(v), pos)
        for v in range(778):
            perm_rev[v] = first[v]
        perm_rev[778:] = rng.randint(0, 1008, size=1008 - 778)
        out[side + '_graph'] = {'coarsen_graphs_L': Ls, 'graph_perm': [int(v) for v in perm], 'graph_perm_reverse': perm_rev}
    out['dense_coor'] = rng.rand(778, 3).astype(np.float32)
    out['upsample'] = (rng.randn(778, 252) * 0.05).astype(np.float32)
    return out


def synthetic_mano(seed=0, side='right'):
    """Random MANO-shaped tensors (never derived from the licensed MANO files)."""
    rng = np.random.RandomState(seed + (0 if side == 'right' else 1000))
    w = rng.rand(778, 16).astype(np.float32) ** 4
    w /= w.sum(1, keepdims=True)
    jreg = np.zeros((16, 778), np.float32)
    for j in range(16):
        idx = rng.choice(778, 12, replace=False)
        v = rng.rand(12).astype(np.float32); jreg[j, idx] = v / v.sum()
    q, _ = np.linalg.qr(rng.randn(45, 45))
    return {
        'hands_components': q.astype(np.float32),
        'hands_mean': (rng.randn(45) * 0.2).astype(np.float32),
        'shapedirs': (rng.randn(778, 3, 10) * 0.005).astype(np.float32),
        'posedirs': (rng.randn(778, 3, 135) * 0.002).astype(np.float32),
        'v_template': (rng.randn(778, 3) * 0.05).astype(np.float32),
        'J_regressor': sp.csc_matrix(jreg),
        'J': (rng.randn(16, 3) * 0.05).astype(np.float32),
        'weights': w,
        'f': np.stack([rng.permutation(778)[:3] for _ in range(1538)]).astype(np.uint32),
        'kintree_table': np.array([[4294967295, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14], list(range(16))], dtype=np.int64),
    }
