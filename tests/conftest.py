import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


@pytest.fixture(scope='session')
def built_lib():
    """Build (if stale) and load the C-ABI library."""
    from hrfuser_b200 import _lib, build
    build.build()
    return _lib.load()


def pytest_sessionfinish(session, exitstatus):
    """HRF_PARITY_LOG=<file>: append the measured parity errors of this run (worst first)."""
    path = os.environ.get('HRF_PARITY_LOG')
    if not path:
        return
    import json
    from helpers import PARITY_LOG
    if PARITY_LOG:
        rows = sorted(PARITY_LOG, key=lambda r: -r[2])
        with open(path, 'a') as f:
            for what, mode, e, out in rows:
                f.write(json.dumps(dict(what=what, mode=mode, err=e, frac_out=out)) + '\n')
