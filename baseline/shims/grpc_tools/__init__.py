"""Stand-in for ``grpcio-tools`` (not installable offline): a minimal proto3 compiler, see ``protoc.py``."""
