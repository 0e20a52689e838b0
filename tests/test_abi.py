"""The C-ABI shared library: loads without a GPU, exports every symbol include/sfx.h declares,
the ctypes mirrors have the sizes the C structs have, and compute entry points fail loudly
(no CPU fallback) when no CUDA device is visible."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from tests import common as Cm
from smplifyx_b200 import _native as N

ROOT = Cm.ROOT


def _declared():
    with open(os.path.join(ROOT, 'include', 'sfx.h')) as f:
        src = f.read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(sfx_[a-z_0-9]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = N.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert lib.sfx_version() >= 100


def test_struct_sizes_match_the_c_side(tmp_path):
    src = tmp_path / 'sz.cpp'
    src.write_text('#include <cstdio>\n#include "%s/include/sfx.h"\nint main(){printf("%%zu %%zu %%zu %%zu\\n",'
                   'sizeof(SfxLayout),sizeof(SfxStage),sizeof(sfx_model_desc),sizeof(SfxPipeline));}\n' % ROOT)
    exe = tmp_path / 'sz'
    subprocess.check_call(['g++', '-std=c++17', '-o', str(exe), str(src)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(N.SfxLayout), C.sizeof(N.SfxStage), C.sizeof(N.SfxModelDesc),
                     C.sizeof(N.SfxPipeline)]


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback():
    lib = N.load_library()
    desc, keep = N.build_model_desc(Cm.model_data(), Cm.joint_map(), **Cm.MODEL_KW)
    h = C.c_void_p()
    rc = lib.sfx_model_create(C.byref(desc), C.byref(h))
    assert rc != 0 and not h.value
    assert b'no CUDA device' in lib.sfx_last_error()
    from smplifyx_b200 import engine
    with pytest.raises(RuntimeError):
        engine.Model(Cm.model_data(), Cm.joint_map(), **Cm.MODEL_KW)


def test_layout_mirror():
    L = N.make_layout(10, 10, 12, False)
    assert L.np == 10 + 3 + 12 + 12 + 3 + 3 + 3 + 10 + 63 + 3
    assert N.make_layout(10, 10, 12, True).n_pose == 32
    st = N.make_stage(L, N.BODY_STAGE_BLOCKS)
    assert st.n_active == L.np - 3 and st.need_blend_grad == 1
    assert st.max_iter == 30 and st.max_eval == 37 and st.history == 100
    cam = N.make_stage(L, N.CAMERA_STAGE_BLOCKS, loss_kind=N.LOSS_CAMERA_INIT)
    assert cam.n_active == 6 and cam.need_blend_grad == 0
