import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on a B200 via gpurun)')
    config.addinivalue_line('markers', 'slow: long-running CPU test')


@pytest.fixture
def workdir(tmp_path, monkeypatch):
    """Isolated RAVNEST_HOME so node_data/ trained/ losses.txt never touch the repo."""
    monkeypatch.setenv('RAVNEST_HOME', str(tmp_path))
    monkeypatch.chdir(tmp_path)
    return tmp_path


_port = [int(os.environ.get('RAVNEST_TEST_PORT_BASE', 21000)) + (os.getpid() % 500) * 16]


@pytest.fixture
def free_ports():
    def take(n):
        base = _port[0]
        _port[0] += n
        return list(range(base, base + n))
    return take
