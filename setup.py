"""pip-installable package; the CUDA extension is built in-tree by `python -m ravnest_b200.ops.build`
(or `__graft_entry__.build()`), not at install time — unlike the reference's setup.py:16-57 which pip-installs
grpcio-tools and compiles protos inside every setuptools command."""
from setuptools import find_packages, setup

setup(name='ravnest_b200', version='0.1.0', packages=find_packages(include=['ravnest_b200', 'ravnest_b200.*']),
      package_data={'ravnest_b200': ['protos/*.proto', 'ops/csrc/*', 'ops/*.so']},
      install_requires=['torch', 'grpcio', 'protobuf', 'numpy'], python_requires='>=3.10',
      description='Blackwell-native asynchronous pipeline + data-parallel training engine with the ravnest API')
