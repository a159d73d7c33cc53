"""GPU: the driver's entry point.  `__graft_entry__.smoke()` runs on a fresh box at round end; this
keeps it inside the GPU suite so a change that breaks it is seen with the other tests."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_smoke_entry_point(capsys):
    import __graft_entry__ as entry
    entry.smoke()
    assert "smoke ok" in capsys.readouterr().out
