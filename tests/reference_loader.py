"""
Test tooling: import the UNMODIFIED reference from /root/reference when it is mounted (this container only;
the GPU box does not have it).  pysam is not installable here, and the reference imports it at module top
level only for type annotations on this path, so a stub module is registered first.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path('/root/reference')


def reference_available() -> bool:
    return (REFERENCE_ROOT / 'demuxalot' / 'demux.py').exists()


def load_reference():
    """Returns the reference `demuxalot` package, or None when /root/reference is absent."""
    if not reference_available():
        return None
    if 'pysam' not in sys.modules:
        stub = types.ModuleType('pysam')
        stub.AlignedRead = type('AlignedRead', (), {})
        stub.AlignedSegment = stub.AlignedRead
        sys.modules['pysam'] = stub
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    import demuxalot  # noqa: E402  (the reference, not demuxalot_b200)
    return demuxalot
