"""
oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU (numpy) restatement of demuxalot's likelihood / EM path, used as the checker for the
CUDA implementation in `demuxalot_b200`.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this package; the product
path (`demuxalot_b200`) never does and fails loudly when its CUDA library is missing.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the unmodified reference from
/root/reference (v0.4.3, with a stub `pysam` module) and stores its inputs/outputs as fixtures;
`tests/test_oracle_golden.py` checks this restatement against every fixture, and
`tests/test_oracle_vs_reference.py` re-runs the comparison live whenever /root/reference exists.
"""
from .demux_oracle import (  # noqa: F401
    OracleDemultiplexer,
    doublet_penalties,
    option_names,
    snp_ids_for_variants,
    match_and_flatten_calls,
    group_molecule_calls,
    regularised_betas,
    probs_from_betas,
    barcode_logits,
    softmax_rows,
    m_step,
    snp_groups,
    snp_aggregated_logits,
    log_softmax_rows,
)
