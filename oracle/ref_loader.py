"""TEST INFRASTRUCTURE ONLY.  Import the *unmodified* reference from /root/reference.

Only usable in the dev container (the GPU box has no /root/reference).  Used by
``oracle/make_golden.py`` to generate the committed fixtures under tests/golden/
and by ``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent).

The reference star-imports utils.py, which needs third-party packages that are not
installed (alphafold2_pytorch/utils.py:12,18-21,24; alphafold2.py:19-20).  They are
never touched by the distogram path, so they are replaced by MagicMock stubs.
"""
import importlib.util
import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("AF2_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "Bio", "Bio.SeqIO", "sidechainnet", "sidechainnet.utils", "sidechainnet.utils.sequence",
    "sidechainnet.utils.measure", "sidechainnet.structure", "sidechainnet.structure.build_info",
    "sidechainnet.structure.StructureBuilder", "mp_nerf", "invariant_point_attention",
    "pytorch3d", "pytorch3d.transforms",
]


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "alphafold2_pytorch", "alphafold2.py"))


def load_reference():
    """Returns the reference module ``alphafold2_pytorch.alphafold2`` (unmodified)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = MagicMock()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    warnings.filterwarnings("ignore", message=".*use_reentrant.*")
    import alphafold2_pytorch.alphafold2 as ref  # noqa
    return ref


def load_reference_rotary():
    """rotary.py is dead code at HEAD (never imported); load it standalone by path."""
    path = os.path.join(REFERENCE_ROOT, "alphafold2_pytorch", "rotary.py")
    spec = importlib.util.spec_from_file_location("_af2_ref_rotary", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
