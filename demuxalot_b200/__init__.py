"""
demuxalot_b200 -- B200-native (sm_100a) likelihood / EM core for demuxalot.

Drop-in for the hot path of arogozhnikov/demuxalot: `Demultiplexer.predict_posteriors` and
`Demultiplexer.learn_genotypes` applied to `count_snps` output, with the same `ProbabilisticGenotypes` /
`BarcodeHandler` / `CompressedSNPCalls` host types and the same betas parquet layout.  The numeric stages are
hand-written CUDA kernels in `csrc/`, reached through the C ABI of `include/demux_b200.h`; there is no CPU
fallback (build the library with `python -m demuxalot_b200.build`).
"""
from .barcodes import BarcodeHandler
from .calls import CompressedSNPCalls
from .counting import count_snps
from .genotype_store import ProbabilisticGenotypes
from .snp_detection import detect_snps_positions

__version__ = '0.1.0'


def __getattr__(name):
    # `Demultiplexer` needs torch; import it lazily so host-only users (I/O, genotype bookkeeping) stay light.
    if name == 'Demultiplexer':
        from .demultiplexer import Demultiplexer
        return Demultiplexer
    raise AttributeError(f'module {__name__!r} has no attribute {name!r}')


__all__ = ['BarcodeHandler', 'CompressedSNPCalls', 'Demultiplexer', 'ProbabilisticGenotypes', 'count_snps',
           'detect_snps_positions']
