"""CPU: the C-ABI library builds, loads and exports exactly what include/gpemsr_b200.h declares."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, 'include', 'gpemsr_b200.h')).read()
    return sorted(set(re.findall(r'GPEMSR_API[^;(]*?\b(gpemsr_\w+)\s*\(', txt)))


def test_header_symbols_exported():
    from gpemsr_b200 import _lib, build
    build.build()
    out = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if ' T ' in l}
    decl = _declared()
    assert len(decl) >= 10
    missing = [s for s in decl if s not in exported]
    assert not missing, missing
    extra = [s for s in exported if s.startswith('gpemsr_') and s not in decl]
    assert not extra, f'exported but undeclared: {extra}'


def test_ctypes_signatures_cover_header():
    from gpemsr_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    L = _lib.lib()
    for name in _lib.SIGNATURES:
        assert hasattr(L, name), name
    assert L.gpemsr_version() >= 0x000100
    assert L.gpemsr_kernel_launches() == 0
    assert L.gpemsr_vq_workspace_bytes(1000, 512, 1024) > 1000 * 512 * 2


def test_no_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import gpemsr_b200
    with pytest.raises(gpemsr_b200.GpemsrError):
        gpemsr_b200.flow_warp(torch.zeros(1, 1, 4, 4), torch.zeros(1, 4, 4, 2))
    with pytest.raises(gpemsr_b200.GpemsrError):
        gpemsr_b200.vq_lookup(torch.zeros(1, 8, 2, 2), torch.zeros(16, 8))
    # the C entry points themselves refuse to run without an sm_100 device
    from gpemsr_b200 import _lib
    rc = _lib.lib().gpemsr_flow_warp(None, None, 1, 1, 4, 4, 0, 1, None, None)
    assert rc in (-3, -4)
    assert b'device' in _lib.lib().gpemsr_last_error_string().lower() or rc == -4


def test_sass_is_blackwell_native():
    """tcgen05 / bulk-copy / TMEM instructions are really in the binary (B200_PROFILING.md mnemonics)."""
    from gpemsr_b200 import _lib
    sass = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True)
    if sass.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    for mnem in ('UTCHMMA', 'LDTM', 'UBLKCP'):
        assert mnem in sass.stdout, mnem
    assert 'HGMMA' not in sass.stdout


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: importing the whole product package (every host mirror) must not pull it in, and no
    product source may name it in an import."""
    import glob
    import os
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ('import sys; import gpemsr_b200; import gpemsr_b200.indexer, gpemsr_b200.vgg, gpemsr_b200.spynet, gpemsr_b200.dcn, '
            'gpemsr_b200.volume, gpemsr_b200.graph, gpemsr_b200.synth_weights; '
            'assert not [m for m in sys.modules if m == "oracle" or m.startswith("oracle.")], "oracle imported"')
    r = subprocess.run([sys.executable, '-c', code], cwd=root, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    for f in glob.glob(os.path.join(root, 'gpemsr_b200', '*.py')):
        assert not re.search(r'^\s*(from|import)\s+oracle\b', open(f).read(), re.M), f
