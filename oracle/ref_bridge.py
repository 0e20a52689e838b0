"""ORACLE (test infrastructure) -- imports the *unmodified* reference modules from
``/root/reference/smplifyx`` so the oracle port (``oracle/fit_port.py``) can be pinned
against the real thing and golden fixtures can be generated (``tests/golden/make_golden.py``).

``/root/reference`` exists only in the authoring container, never on the GPU box: everything
that calls ``load()`` must be skipped when ``available()`` is False.

Non-arithmetic third-party imports of the reference (open3d, skimage, trimesh, pyrender,
plyfile, human_body_prior, configargparse) are absent here and are stubbed; ``smplx`` is
replaced by the restatement in ``oracle/smplx_shim.py``.  No reference
source is copied: the modules are imported from where they lie.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get('SFX_REFERENCE_ROOT', '/root/reference')
REF_PKG = os.path.join(REF_ROOT, 'smplifyx')

_loaded = None


def available():
    return os.path.isfile(os.path.join(REF_PKG, 'fitting.py'))


class _Anything(types.ModuleType):
    """A module whose every attribute is a harmless callable / sub-stub."""

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        obj = _Anything(self.__name__ + '.' + name)
        setattr(self, name, obj)
        return obj

    def __call__(self, *a, **k):
        return None


class _PlyElement(object):
    @staticmethod
    def describe(data, name):
        return (name, data)


class _PlyData(object):
    def __init__(self, elements, text=False, byte_order='<'):
        self.elements = elements

    def write(self, path):
        # enough for the parity tests: x,y,z float32 little-endian payload
        import numpy as np
        name, data = self.elements[0]
        np.save(path + '.npy', np.stack([data['x'], data['y'], data['z']], axis=1))


def _install_stubs():
    from oracle import smplx_shim
    for name in ['open3d', 'skimage', 'skimage.io', 'skimage.transform', 'trimesh',
                 'pyrender', 'configargparse', 'human_body_prior',
                 'human_body_prior.tools', 'human_body_prior.tools.model_loader',
                 'human_body_prior.tools.visualization_tools',
                 'human_body_prior.body_model', 'human_body_prior.body_model.body_model']:
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    ply = types.ModuleType('plyfile')
    ply.PlyElement = _PlyElement
    ply.PlyData = _PlyData
    sys.modules['plyfile'] = ply
    smplx = types.ModuleType('smplx')
    smplx.create = smplx_shim.create
    lbs = types.ModuleType('smplx.lbs')
    lbs.transform_mat = smplx_shim.transform_mat
    smplx.lbs = lbs
    sys.modules['smplx'] = smplx
    sys.modules['smplx.lbs'] = lbs


def load():
    """Returns a namespace with the reference modules: fitting, camera, prior, utils,
    optimizers (optim_factory, lbfgs_ls), fit_single_frame, data_parser."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference tree not present at ' + REF_ROOT)
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if here not in sys.path:
        sys.path.insert(0, here)
    _install_stubs()
    # the reference uses flat imports ("import utils"): put its directory on sys.path under
    # guard so that our own modules of the same name are not shadowed elsewhere.
    saved = {k: sys.modules.get(k) for k in
             ['utils', 'fitting', 'camera', 'prior', 'optimizers', 'mesh_viewer',
              'data_parser', 'fit_single_frame', 'optimizers.optim_factory',
              'optimizers.lbfgs_ls']}
    for k in saved:
        sys.modules.pop(k, None)
    sys.path.insert(0, REF_PKG)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ns = types.SimpleNamespace()
            ns.utils = importlib.import_module('utils')
            ns.prior = importlib.import_module('prior')
            ns.camera = importlib.import_module('camera')
            ns.fitting = importlib.import_module('fitting')
            ns.optim_factory = importlib.import_module('optimizers.optim_factory')
            ns.lbfgs_ls = importlib.import_module('optimizers.lbfgs_ls')
            ns.data_parser = importlib.import_module('data_parser')
            ns.fit_single_frame = importlib.import_module('fit_single_frame')
    finally:
        sys.path.remove(REF_PKG)
        # keep reference modules alive under private names, restore the public ones
        for k in list(saved):
            mod = sys.modules.pop(k, None)
            if mod is not None:
                sys.modules['_sfxref_' + k] = mod
            if saved[k] is not None:
                sys.modules[k] = saved[k]
    _loaded = ns
    return ns


def load_cmd_parser():
    """The reference's unmodified ``cmd_parser`` module (smplifyx/cmd_parser.py) running on the
    restated ``configargparse`` (oracle/configargparse_shim.py)."""
    import importlib.util
    from oracle import configargparse_shim
    if not available():
        raise RuntimeError('reference tree not present at ' + REF_ROOT)
    saved = sys.modules.get('configargparse')
    sys.modules['configargparse'] = configargparse_shim
    try:
        spec = importlib.util.spec_from_file_location('_sfxref_cmd_parser',
                                                      os.path.join(REF_PKG, 'cmd_parser.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is not None:
            sys.modules['configargparse'] = saved
        else:
            sys.modules.pop('configargparse', None)
    return mod
