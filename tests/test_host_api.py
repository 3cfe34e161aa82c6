"""Host-side mirror of the reference interface (no GPU): types, layouts, error behaviour, C-ABI exports."""
import re
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from demuxalot_b200 import BarcodeHandler, CompressedSNPCalls, ProbabilisticGenotypes
from demuxalot_b200.calls import MOLECULE_DTYPE, SNP_CALL_DTYPE
from demuxalot_b200.synthetic import make_dataset
from golden_io import GOLDEN_DIR, bits

ROOT = Path(__file__).resolve().parent.parent


def test_abi_library_exports_every_declared_symbol(native_lib):
    """include/demux_b200.h is the contract: every dmx_* function it declares must be exported and typed."""
    from demuxalot_b200 import _native
    header = (ROOT / 'include' / 'demux_b200.h').read_text()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(dmx_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    for name in declared:
        assert hasattr(native_lib, name), name
    assert native_lib.dmx_abi_version() == _native.ABI_VERSION == 4
    # size queries are pure host functions: callable without a GPU
    assert native_lib.dmx_estep_workspace_bytes(10, 4, 0.25, 0, 1) >= 10 * 10 * 4
    assert native_lib.dmx_estep_workspace_bytes(10, 4, 0.0, 0, 1) >= 10 * 4 * 4
    assert native_lib.dmx_estep_workspace_bytes(10, 32, 0.35, 10, 0) == 0  # one item per barcode: no segment sums
    assert native_lib.dmx_estep_workspace_bytes(10, 32, 0.35, 14, 0) >= 14 * 528 * 8
    # widths served by the warp-per-item pair kernel (FAST flavour, doublet columns)
    def blocks_ok(nb):  # widths of the patch kernel: at least 65 % of the lanes of its ceil(T / 32) warps carry tiles
        tiles = nb * (nb + 1) // 2
        return 9 <= nb <= 32 and 100 * tiles >= 65 * 32 * -(-tiles // 32)
    # (g + 7) // 8 == 4 and == 8: the strip kernel (estep_pairs_strip.cu)
    want = [g for g in range(1, 270) if g <= 16 or (g + 7) // 8 in (3, 4, 5, 7, 8) or blocks_ok((g + 7) // 8)]
    assert [g for g in range(1, 270) if native_lib.dmx_estep_plan_supported(g, 0.35, 1)] == want
    assert 4 in want and 32 in want and 200 in want and 100 in want and 72 in want and 88 in want and 64 in want and 16 in want and 44 not in want
    assert not native_lib.dmx_estep_plan_supported(32, 0.0, 1) and not native_lib.dmx_estep_plan_supported(32, 0.35, 0)
    assert native_lib.dmx_estep_plan_supported(4, 0.0, 1) and not native_lib.dmx_estep_plan_supported(9, 0.0, 1)


def test_ctypes_signatures_match_the_header_prototypes():
    """Every parameter of every prototype in include/demux_b200.h against the ctypes binding: same count, pointers
    bound as pointers, integers / floats with the declared width (a mismatch would corrupt the call silently)."""
    import ctypes as C
    from demuxalot_b200 import _native
    header = (ROOT / 'include' / 'demux_b200.h').read_text()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    scalar = {'int32_t': C.c_int32, 'int64_t': C.c_int64, 'int': C.c_int, 'float': C.c_float, 'double': C.c_double}
    prototypes = re.findall(r'([A-Za-z_][A-Za-z0-9_ ]*?[ *]+)\b(dmx_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', header)
    assert len(prototypes) == len(_native.SIGNATURES)
    for restype_text, name, params in prototypes:
        restype, argtypes = _native.SIGNATURES[name]
        restype_text = restype_text.strip()
        if restype_text.endswith('*'):
            assert restype in (C.c_char_p, C.c_void_p), name
        else:
            assert restype is scalar[restype_text], (name, restype_text)
        params = [p.strip() for p in params.split(',')] if params.strip() not in ('', 'void') else []
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        for text, bound in zip(params, argtypes):
            if '*' in text:
                assert bound is C.c_void_p or issubclass(bound, C._Pointer), (name, text, bound)
            else:
                kind = text.replace('const ', '').split()[0]
                assert bound is scalar[kind], (name, text, bound)


def test_host_gather_of_the_barcode_column(native_lib):
    """dmx_host_gather_cb is a host function (no GPU needed): the compressed_cb column of packed molecule records."""
    import ctypes as C
    from demuxalot_b200.calls import MOLECULE_DTYPE
    rng = np.random.default_rng(0)
    for n, threads in ((0, 4), (1, 1), (70_001, 1), (300_000, 7), (300_000, 64)):
        mols = np.zeros(n, dtype=MOLECULE_DTYPE)
        assert mols.dtype.itemsize == 12
        mols['compressed_cb'] = rng.integers(-5, 2 ** 31 - 1, size=n)
        mols['compressed_ub'] = rng.integers(0, 2 ** 31 - 1, size=n)
        mols['p_group_misaligned'] = rng.random(n)
        out = np.full(n + 3, -77, dtype=np.int32)
        rc = native_lib.dmx_host_gather_cb(C.c_void_p(mols.ctypes.data), n, C.c_void_p(out.ctypes.data), threads)
        assert rc == 0
        assert np.array_equal(out[:n], mols['compressed_cb']) and (out[n:] == -77).all()


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from demuxalot_b200 import Demultiplexer
    ds = make_dataset(n_genotypes=2, n_snps=20, n_barcodes=4, rows_per_barcode=5, seed=1)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        Demultiplexer.predict_posteriors(ds.calls, ds.genotypes, ds.barcode_handler)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        Demultiplexer.learn_genotypes(ds.calls, ds.genotypes, ds.barcode_handler)


def test_product_does_not_import_oracle():
    for path in (ROOT / 'demuxalot_b200').glob('*.py'):
        text = path.read_text()
        assert 'import oracle' not in text and 'from oracle' not in text, path


def test_barcode_handler():
    bh = BarcodeHandler(['TTT-1', 'AAA-1', 'CCC-1'])
    assert bh.ordered_barcodes == ['AAA-1', 'CCC-1', 'TTT-1'] and bh.n_barcodes == 3
    assert bh.barcode2index['CCC-1'] == 1

    class Read:
        def __init__(self, **tags): self.tags = tags
        def has_tag(self, t): return t in self.tags
        def get_tag(self, t): return self.tags[t]
    assert bh.get_barcode_index(Read(CB='TTT-1')) == 2
    assert bh.get_barcode_index(Read(CB='GGG-1')) is None
    assert bh.get_barcode_index(Read(UB='x')) is None
    with pytest.raises(AssertionError):
        BarcodeHandler('barcodes.tsv')
    with pytest.raises(AssertionError):
        BarcodeHandler(['A', 'A'])
    rg = BarcodeHandler(['A', 'A', 'B'], RG_tags=['x', 'y', 'x'])
    assert rg.ordered_barcodes == [('A', 'x'), ('A', 'y'), ('B', 'x')]
    assert rg.get_barcode_index(Read(CB='A', RG='y')) == 1
    sub = rg.filter_to_rg_value('x')
    assert sub.n_barcodes == 3 and sub.get_barcode_index(Read(CB='B')) == 2 and sub.get_barcode_index(Read(CB='A')) == 0


def test_barcode_handler_from_file(tmp_path):
    path = tmp_path / 'barcodes.csv'
    path.write_text('GGG-1\nAAA-1\n')
    assert BarcodeHandler.from_file(path).ordered_barcodes == ['AAA-1', 'GGG-1']


def test_compressed_snp_calls_layout_and_growth():
    assert SNP_CALL_DTYPE.itemsize == 13 and MOLECULE_DTYPE.itemsize == 12
    c = CompressedSNPCalls(start_snps_size=2, start_molecule_size=1)
    for m in range(5):
        c.add_calls_from_read_group(m, 100 + m, 0.01, [(10 * m + k, 'ACGTN'[k % 5], 0.001) for k in range(3)])
    assert c.n_molecules == 5 and c.n_snp_calls == 15
    assert c.snp_calls.dtype == SNP_CALL_DTYPE and c.molecules.dtype == MOLECULE_DTYPE
    assert list(c.snp_calls['base_index'][:5]) == [0, 1, 2, 0, 1]
    assert list(c.snp_calls['molecule_index'][:15:3]) == [0, 1, 2, 3, 4]
    merged = CompressedSNPCalls.concatenate([c, c])
    assert merged.n_molecules == 10 and merged.n_snp_calls == 30
    assert merged.snp_calls['molecule_index'][15] == 5
    assert merged.snp_calls.dtype == SNP_CALL_DTYPE and merged.molecules.dtype == MOLECULE_DTYPE
    assert np.array_equal(merged.molecules, np.concatenate([c.molecules[:5], c.molecules[:5]]))
    shifted = c.snp_calls[:15].copy()
    shifted['molecule_index'] += 5
    assert np.array_equal(merged.snp_calls, np.concatenate([c.snp_calls[:15], shifted]))
    assert c.snp_calls['molecule_index'][0] == 0  # the parts are left untouched
    empty = CompressedSNPCalls.concatenate([])
    assert empty.n_molecules == 0 and empty.n_snp_calls == 0
    c.minimize_memory_footprint()
    assert len(c.snp_calls) == 15 and len(c.molecules) == 5


def test_genotypes_store_basics_and_index():
    g = ProbabilisticGenotypes(['D1', 'D2'])
    with pytest.raises(AssertionError):
        ProbabilisticGenotypes(['D2', 'D1'])
    a = g.get_variant_id('chr1', 10, 'A')
    b = g.get_variant_id('chr2', 5, 'C')
    c = g.get_variant_id('chr1', 10, 'G')
    assert (a, b, c) == (0, 1, 2) and g.get_variant_id('chr1', 10, 'A') == 0
    g.variant_betas[:3] = [[1, 2], [3, 4], [5, 6]]
    assert g.n_variants == 3 and g.get_betas().shape == (3, 2) and not g.get_betas().flags.writeable
    assert list(g.get_snp_ids_for_variants()) == [0, 1, 0]  # first-seen order of (chrom, pos)
    idx = g.hot_path_index()
    assert list(idx['snp_offsets']) == [0, 2, 3] and list(idx['snp_variants']) == [0, 2, 1]
    assert np.all(np.diff(idx['keys_sorted']) > 0)
    assert g.hot_path_index() is idx  # cached
    g.get_variant_id('chr3', 1, 'T')
    assert g.hot_path_index() is not idx  # invalidated by growth
    learnt = g._with_betas(np.ones((4, 2), dtype=np.float32))
    assert learnt is not g and learnt.variant_betas.shape == (4, 2) and g.variant_betas[0, 0] == 1
    with pytest.raises(AssertionError):
        g._with_betas(-np.ones((4, 2), dtype=np.float32))
    with pytest.raises(AssertionError):
        g._with_betas(np.ones((4, 2), dtype=np.float64))
    pos = g.get_chromosome2positions()
    assert list(pos['chr1']) == [10] and set(pos) == {'chr1', 'chr2', 'chr3'}
    # capacity doubling keeps contents (genotypes.py:75-78)
    for k in range(40000):
        g.get_variant_id('chrX', k, 'A')
    assert g.n_variants == 40004 and len(g.variant_betas) >= 40004 and g.variant_betas[2, 1] == 6


def test_betas_parquet_layout_and_roundtrip(tmp_path):
    ds = make_dataset(n_genotypes=3, n_snps=20, n_barcodes=8, rows_per_barcode=5, seed=105)  # as in make_golden.py
    g = ds.genotypes
    path = tmp_path / 'betas.parquet'
    g.save_betas(path)
    import pyarrow.parquet as pq
    ours, golden = pq.read_table(path), pq.read_table(GOLDEN_DIR / 'reference_betas.parquet')
    assert ours.schema.equals(golden.schema), (ours.schema, golden.schema)
    assert ours.equals(golden)  # same rows in the same (chrom, pos, base) order, float32 columns
    frame = pd.read_parquet(path)
    assert list(frame.index.names) == ['CHROM', 'POS', 'BASE'] and list(frame.columns) == g.genotype_names
    assert all(frame.dtypes == np.float32)
    # reference tests/test_synthetic.py:241-260
    g2 = ProbabilisticGenotypes(g.genotype_names, default_prior=g.default_prior)
    g2.add_prior_betas(GOLDEN_DIR / 'reference_betas.parquet')
    assert set(g.var2varid) == set(g2.var2varid)
    for variant, vid in g.var2varid.items():
        assert np.allclose(g.variant_betas[vid], g2.variant_betas[g2.var2varid[variant]])
    g2.add_prior_betas(path, prior_strength=2.)  # accumulates
    for variant, vid in g.var2varid.items():
        assert np.allclose(3 * g.variant_betas[vid], g2.variant_betas[g2.var2varid[variant]])


def test_add_vcf_plain_text(tmp_path):
    vcf = tmp_path / 'g.vcf'
    vcf.write_text(
        '##fileformat=VCFv4.2\n'
        '#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tD1\tD2\tD3\tEXTRA\n'
        'chr1\t100\t.\tA\tG\t.\t.\t.\tGT\t0/0\t0/1\t1/1\t0/0\n'
        'chr1\t200\t.\tC\tT\t.\t.\t.\tGT:DP\t0|1:3\t./.:0\t1/1:9\t0/0:1\n'
        'chr1\t300\t.\tAT\tG\t.\t.\t.\tGT\t0/0\t0/1\t1/1\t0/0\n'   # not a SNV
        'chr2\t50\t.\tG\tA\t.\t.\t.\tGT\t0/0\t./.\t./.\t0/0\n'     # fewer than two called donors
    )
    g = ProbabilisticGenotypes(['D1', 'D2', 'D3'])
    with pytest.warns(UserWarning):
        g.add_vcf(vcf)
    assert set(g.var2varid) == {('chr1', 99, 'A'), ('chr1', 99, 'G'), ('chr1', 199, 'C'), ('chr1', 199, 'T'),
                                ('chr2', 49, 'G'), ('chr2', 49, 'A')}
    b = g.variant_betas
    assert list(b[g.var2varid['chr1', 99, 'A']]) == [100, 50, 0]
    assert list(b[g.var2varid['chr1', 99, 'G']]) == [0, 50, 100]
    # D2 not called at chr1:200 -> 0.1 x mean of the called donors per allele
    assert np.allclose(b[g.var2varid['chr1', 199, 'C']], [50, 0.1 * 25, 0])
    assert np.allclose(b[g.var2varid['chr1', 199, 'T']], [50, 0.1 * 75, 100])
    assert np.all(b[g.var2varid['chr2', 49, 'G']] == 0)  # registered, but the record was skipped


def test_synthetic_generator_is_deterministic_and_adversarial():
    a = make_dataset(n_genotypes=5, n_snps=200, n_barcodes=30, rows_per_barcode=60, seed=3, shuffle_variants=True,
                     spare_capacity=5)
    b = make_dataset(n_genotypes=5, n_snps=200, n_barcodes=30, rows_per_barcode=60, seed=3, shuffle_variants=True,
                     spare_capacity=5)
    assert a.genotypes.var2varid == b.genotypes.var2varid
    assert np.array_equal(a.genotypes.variant_betas, b.genotypes.variant_betas)
    for chrom in a.calls:
        assert np.array_equal(a.calls[chrom].snp_calls, b.calls[chrom].snp_calls)
        assert len(a.calls[chrom].snp_calls) == a.calls[chrom].n_snp_calls + 5  # over-allocated
        assert a.calls[chrom].n_molecules < a.calls[chrom].n_snp_calls  # some molecules carry two calls
    v2s = a.genotypes.get_snp_ids_for_variants()
    assert (np.bincount(v2s) == 3).any()  # multi-allelic positions exist
    import oracle
    _, _, mol, rows = oracle.OracleDemultiplexer.pack_calls(a.calls, a.genotypes, True)
    assert len(mol['variant_id']) < a.n_calls  # unmatched calls exist
    assert rows['barcode_variant_count'].max() > 1  # UMI-combination happens
    assert len(np.unique(rows['compressed_cb'])) < 30 or True


def test_demultiplexer_host_helpers_match_oracle():
    import oracle
    from demuxalot_b200.demultiplexer import Demultiplexer, option_names, n_options
    for g in (1, 2, 3, 10, 32):
        for dp in (0., 0.25, 0.35, 0.5):
            assert np.array_equal(bits(Demultiplexer._doublet_penalties(g, dp)), bits(oracle.doublet_penalties(g, dp)))
            names = [f'D{k:02d}' for k in range(g)]
            assert option_names(names, dp) == oracle.option_names(names, dp)
            assert len(option_names(names, dp)) == n_options(g, dp)
    assert Demultiplexer.contribution_power == 2. and Demultiplexer.aggregate_on_snps is False


@pytest.mark.parametrize('n_blocks', [4, 8])
def test_strip_kernel_layout_covers_every_pair_exactly_once(native_lib, n_blocks):
    """Host-side check of csrc/estep_pairs_strip.cu::strip_layout (no GPU): the 36 packed product slots of every lane
    slot, decoded as the kernel's epilogue decodes them (strip_pair_of), produce every genotype pair i <= j of the padded
    width exactly once; strips 0, 1 share a block, and the operand loads are aligned as the kernel assumes."""
    import ctypes as C
    raw = (C.c_int16 * (32 * 8))()
    n_slots = C.c_int32(0)
    assert native_lib.dmx_estep_strip_layout(n_blocks, raw, C.byref(n_slots)) == 0
    assert n_slots.value == (8 if n_blocks == 4 else 32)
    G = 8 * n_blocks
    seen = {}
    for s in range(n_slots.value):
        p0, p1, p2, q1, d, P1, P2, _ = raw[8 * s:8 * s + 8]
        for offset in (p0, p1, p2, P1):
            assert 0 <= offset < G and offset % 2 == 0  # LDS.64 of a pair
        assert q1 % 8 == 0 and d % 8 == 0 and 0 <= q1 < G and 0 <= d < G  # LDS.128 of a block
        assert P2 < 0 or (P2 % 2 == 0 and P2 // 8 == d // 8)
        for q in range(36):
            for el in (0, 1):
                if q < 16:
                    pair, other = (p0 if q < 8 else p1), q1 + (q & 7)
                elif q < 24:
                    pair, other = p2, d + (q & 7)
                elif q < 32:
                    pair, other = P1, d + (q & 7)
                else:
                    pair, other = P2, d + 4 + (q & 3)
                if pair < 0:
                    continue
                x = pair + el
                if q >= 24 and pair // 8 == d // 8 and other < x:
                    continue  # lower triangle of a diagonal block
                key = (min(x, other), max(x, other))
                assert key not in seen, (key, seen[key], (s, q, el))
                seen[key] = (s, q, el)
    assert set(seen) == {(i, j) for i in range(G) for j in range(i, G)}
    if n_blocks == 4:  # every slot's unit 3 is a diagonal half whose pairs sit at (2h, 2h+1) / (6-2h, 7-2h) of block d
        for s in range(8):
            _, _, _, _, d, P1, P2, _ = raw[8 * s:8 * s + 8]
            h = (P1 - d) // 2
            assert h in (0, 1) and P2 - d == (4 if h else 6)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) on a shrunken workload: one JSON line with the
    keys of the bench contract, `value` = resident step, `e2e` = the same plus the pack."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--scale', '0.02', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [line for line in out.stdout.splitlines() if line.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
                'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['gpu_launches'] == 0 and line['vs_baseline'] is None
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and 0 < line['e2e']['value'] <= line['value']
    assert 'workload' in line['config']
