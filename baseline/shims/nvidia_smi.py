"""Stand-in for ``nvidia-ml-py3``'s ``nvidia_smi`` module: same NVML entry points, served by ``pynvml``
(``nvidia-ml-py`` is installed; the reference calls nvmlInit/DeviceGetCount/…/Shutdown, ``ravnest/utils.py:211-221``)."""
from pynvml import *  # noqa: F401,F403
