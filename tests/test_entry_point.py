"""CPU (authoring container only): the reference's REAL ``output_GPEMSR.main()`` runs unmodified on the mirror through
``gpemsr_b200.dropin.install()`` -- yml, dataset, constructor, checkpoint loads with the reference's full key set, the padded
head / tail windows and the loader loop, tensor2img, PNG writes (tools/run_entry_point.py; the C library is the argument-checking
stand-in, so no pixels are computed here -- the numerical half of the same loop is tests/test_entry_loop_gpu.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/GPEMSR-CREMI/GPEMSR'


@pytest.mark.skipif(not os.path.isdir(REF), reason='needs the reference tree (authoring container)')
@pytest.mark.parametrize('scale', [8, 16])
def test_real_entry_point_runs_on_the_mirror(scale):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'run_entry_point.py'), '--scale', str(scale)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    assert f'ran unmodified on gpemsr_b200.GPEMSR (x{scale}): 9 windows' in r.stdout
