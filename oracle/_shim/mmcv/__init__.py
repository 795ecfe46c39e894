"""Minimal stand-in for the parts of mmcv the CODD reference imports.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): lets the unmodified reference
files under /root/reference be imported in a container that has no mmcv, so the
oracle restatement can be pinned against them and golden vectors generated.
"""
from .utils import Registry, mkdir_or_exist  # noqa: F401


def is_list_of(seq, expected_type):
    return isinstance(seq, list) and all(isinstance(s, expected_type) for s in seq)
