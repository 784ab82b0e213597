"""Relation proposal networks (lib/modeling/relpn/__init__.py re-exports make_relpn)."""
from .relpn import RelPN, make_relpn  # noqa: F401
from .ppn import PPN, PPNHead, make_ppn  # noqa: F401
from .dpn import DPN, DPNHead, make_dpn  # noqa: F401
from .rel_nms import RelNMS  # noqa: F401
from .anchor_generator import AnchorGenerator, generate_anchors, make_anchor_generator  # noqa: F401
