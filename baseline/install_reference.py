#!/usr/bin/env python
"""Install the UNMODIFIED reference (ravenprotocol/ravnest) into ``baseline/_ref`` for ``bench.py --impl reference``.

Why not ``pip install``: the reference's ``setup.py`` (``:16-21``) pip-installs ``grpcio-tools==1.51.3`` from an index
inside *every* setuptools command to compile its protos; there is no index and no such wheel in ``/opt/wheelhouse``, so
``pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy>``
fails in ``egg_info`` (recorded in DESIGN.md §0).  This script does what that install would have done:

1. copy the source tree verbatim (``ravnest/``, ``models.py``, ``examples/``, ``cluster_formation.py`` …) — no file of
   the reference is edited;
2. run the reference's own proto build step (the body of ``setup.py: proto_compile``: same protoc command line, same
   relative-import rewrite) with ``baseline/shims/grpc_tools`` standing in for the missing compiler.

Run-time dependencies the image lacks (``pippy``, ``torchinfo``, ``nvidia_smi``, ``torch_optimizer``) are provided as
stand-in modules under ``baseline/shims`` and put on ``sys.path`` by the bench harness; they contain no ravnest_b200 code.
"""
import glob
import hashlib
import json
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref')
SHIMS = os.path.join(HERE, 'shims')
SOURCE = os.environ.get('RAVNEST_REFERENCE', '/root/reference')


def tree_digest(root):
    """sha256 over (relative path, content) of every reference source file — proves the copy is unmodified."""
    h = hashlib.sha256()
    for folder, _, files in sorted(os.walk(root)):
        if '__pycache__' in folder:
            continue
        for f in sorted(files):
            if f.endswith(('.py', '.proto')) and not re.search(r'_pb2(_grpc)?\.py$', f):
                path = os.path.join(folder, f)
                h.update(os.path.relpath(path, root).encode())
                with open(path, 'rb') as fh:
                    h.update(fh.read())
    return h.hexdigest()


def installed():
    return os.path.exists(os.path.join(DEST, 'ravnest', 'protos', 'server_pb2_grpc.py'))


def install(force=False):
    if installed() and not force:
        return DEST
    if not os.path.isdir(os.path.join(SOURCE, 'ravnest')):
        raise RuntimeError('reference tree not found at {} and baseline/_ref is not populated'.format(SOURCE))
    if os.path.exists(DEST):
        shutil.rmtree(DEST)
    shutil.copytree(SOURCE, DEST, ignore=shutil.ignore_patterns('.git', '__pycache__'))
    for folder, dirs, files in os.walk(DEST):          # the mount is read-only; the copy must be writable
        for name in dirs + files:
            os.chmod(os.path.join(folder, name), 0o755 if name in dirs else 0o644)
    # --- the reference's proto_compile(), with the stand-in compiler ---------------------------------------
    sys.path.insert(0, SHIMS)
    import grpc_tools.protoc
    cwd = os.getcwd()
    os.chdir(DEST)
    try:
        out = '{}/ravnest/protos'.format(DEST)
        cli_args = ['grpc_tools.protoc', '--proto_path=ravnest/protos', '--python_out=' + out, '--pyi_out=' + out,
                    '--grpc_python_out=' + out] + glob.glob('ravnest/protos/*.proto')
        code = grpc_tools.protoc.main(cli_args)
        assert code == 0
        for script in glob.iglob('{}/ravnest/protos/*.py'.format(DEST)):
            with open(script, 'r+') as fh:
                text = fh.read()
                fh.seek(0)
                fh.write(re.sub(r'\n(import .+_pb2.*)', '\nfrom . \\1', text))
                fh.truncate()
    finally:
        os.chdir(cwd)
    meta = {'source': SOURCE, 'source_digest': tree_digest(SOURCE), 'copy_digest': tree_digest(DEST),
            'generated': sorted(os.path.basename(p) for p in glob.glob(os.path.join(DEST, 'ravnest/protos/*_pb2*.py')))}
    assert meta['source_digest'] == meta['copy_digest'], 'copy differs from the reference tree'
    with open(os.path.join(DEST, 'INSTALL_RECORD.json'), 'w') as fh:
        json.dump(meta, fh, indent=1)
    return DEST


if __name__ == '__main__':
    print(install(force='--force' in sys.argv))
    print(open(os.path.join(DEST, 'INSTALL_RECORD.json')).read())
