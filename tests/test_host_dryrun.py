"""CPU: the host orchestration of gpemsr_b200.GPEMSR (forward, forward_volume) runs end to end against a mocked C library
that only checks argument counts (tools/dryrun_full.py) -- names, shapes, buffer plans and call lists, no compute."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_whole_model_host_logic_dry_run():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'dryrun_full.py')], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = r.stdout
    assert '8 (1, 1, 128, 128) (1, 5, 1, 128, 128)' in out          # x8, 16 x 16 window
    assert '16 (1, 1, 320, 384) (1, 5, 1, 320, 384)' in out         # x16, 20 x 24 window
    assert 'volume (7, 1, 128, 128)' in out and 'volume block (3, 1, 128, 128)' in out
