"""Recipe that stages the UNMODIFIED reference for use on the GPU box:  python -m oracle.build_ref

TEST / BASELINE INFRASTRUCTURE -- never imported by the product package.

The reference (adwardlee/RenderIH) is pure Python, so "building" it means making its own source files importable where
/root/reference does not exist (the GPU box).  This script copies the Python / YAML files of the directories the hot path and its
callers live in (models, utils, core, common, main, dataset, apps) byte for byte from /root/reference into `oracle/_ref/src/`
and unpacks the asset bundle `misc.tar` (a ZIP) into `oracle/_ref/misc/`.  `oracle/_ref/` is git-ignored (no reference source
enters this repository's history) but NOT gpurun-ignored, so it travels with the snapshot like our own built `.so`.
`oracle/ref_bridge.py` imports the reference from /root/reference when it exists and from `oracle/_ref/src` otherwise; a
MANIFEST with the sha256 of every copied file is written so that tests can assert the staged copy is unmodified.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('RIH_REFERENCE_ROOT', '/root/reference')
DST = os.path.join(HERE, '_ref', 'src')
DIRS = ('models', 'utils', 'core', 'common', 'main', 'dataset', 'apps')
KEEP = ('.py', '.yaml', '.yml')


def sha256(path):
    h = hashlib.sha256()
    with open(path, 'rb') as f:
        h.update(f.read())
    return h.hexdigest()


def stage(verbose=True):
    if not os.path.isdir(os.path.join(SRC, 'models')):
        raise RuntimeError('reference tree not found at %s' % SRC)
    manifest = {}
    for d in DIRS:
        for root, _dirs, files in os.walk(os.path.join(SRC, d)):
            for fn in files:
                if not fn.endswith(KEEP):
                    continue
                src = os.path.join(root, fn)
                rel = os.path.relpath(src, SRC)
                dst = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                manifest[rel] = sha256(dst)
    commit = None
    try:
        with open(os.path.join(SRC, '.SUBMODULES.json')) as f:
            commit = json.load(f)
    except Exception:
        pass
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': SRC, 'commit_info': commit, 'files': manifest}, f, indent=1, sort_keys=True)
    from oracle import ref_bridge
    ref_bridge.extract_assets()
    if verbose:
        print('staged %d reference files under %s' % (len(manifest), DST))
    return DST


def verify():
    """-> list of staged files whose content no longer matches the manifest (empty = unmodified copy)."""
    with open(os.path.join(DST, 'MANIFEST.json')) as f:
        files = json.load(f)['files']
    return [rel for rel, h in files.items() if not os.path.exists(os.path.join(DST, rel)) or sha256(os.path.join(DST, rel)) != h]


if __name__ == '__main__':
    sys.path.insert(0, os.path.dirname(HERE))
    stage()
