"""Stage the unmodified reference under ``baseline/_ref/`` (git-ignored, NOT gpurun-ignored: it travels to the
GPU box, where ``/root/reference`` does not exist).

The reference is a flat script tree without ``setup.py`` / ``pyproject.toml``, so
``pip install --target baseline/_ref /root/reference`` has nothing to install; the packages the detect / loss path
imports — ``layers/``, ``data/`` (for ``data.config``), ``utils/`` (for ``utils/nms/py_cpu_nms.py``) — are copied
byte for byte instead.  ``MANIFEST.json`` records the sha256 of every copied file next to the source path so that
"unmodified" can be checked.  Called by ``__graft_entry__.build()`` when the checkout is present."""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get('RD_REFERENCE', '/root/reference')
REF_DST = os.path.join(HERE, '_ref')
PACKAGES = ('layers', 'data', 'utils')
KEEP_EXT = ('.py',)


def build_ref(force=False):
    """Returns the staging directory, or None when the reference checkout is absent (GPU box: prebuilt copy)."""
    if not os.path.isdir(os.path.join(REF_SRC, 'layers')):
        return REF_DST if os.path.isdir(os.path.join(REF_DST, 'layers')) else None
    manifest = {}
    for pkg in PACKAGES:
        for root, dirs, files in os.walk(os.path.join(REF_SRC, pkg)):
            dirs[:] = [d for d in dirs if d != '__pycache__']
            for name in files:
                if not name.endswith(KEEP_EXT):
                    continue
                src = os.path.join(root, name)
                rel = os.path.relpath(src, REF_SRC)
                dst = os.path.join(REF_DST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                with open(src, 'rb') as f:
                    data = f.read()
                digest = hashlib.sha256(data).hexdigest()
                manifest[rel] = digest
                if force or not os.path.exists(dst) or hashlib.sha256(open(dst, 'rb').read()).hexdigest() != digest:
                    shutil.copyfile(src, dst)
    with open(os.path.join(REF_DST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': REF_SRC, 'sha256': manifest}, f, indent=1, sort_keys=True)
    return REF_DST


if __name__ == '__main__':
    print(build_ref(force=True))
