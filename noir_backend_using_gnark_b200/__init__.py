"""noir_backend_using_gnark_b200 — the sm_100a PLONK prover hot path (BN254 G1 MSM + fr NTT) behind the
reference's arithmetic interface.  See DESIGN.md; the C ABI is include/b200zk.h."""
from ._lib import DIF, DIT, B200zkError, header_symbols, load  # noqa: F401
from .api import BitReverse, Commit, Context, Domain, MultiExp, MultiExpShard, SRS, SumPartials, default_context  # noqa: F401

__all__ = [
    "DIF", "DIT", "B200zkError", "Context", "Domain", "BitReverse", "SRS", "MultiExp", "MultiExpShard", "Commit", "SumPartials",
    "default_context", "load", "header_symbols",
]
