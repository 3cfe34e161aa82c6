"""
`ProbabilisticGenotypes` -- per-donor Dirichlet weights ("betas") for every known variant.

Host-side mirror of the reference's class (demuxalot/genotypes.py:18-361): same attributes
(`var2varid`, `variant_betas`, `genotype_names`, `default_prior`), same accessors used by the hot path
(`get_betas`, `get_snp_ids_for_variants`, `_with_betas`, `n_variants`, `n_genotypes`) and the same betas
parquet layout (MultiIndex CHROM/POS/BASE, one float32 column per donor; genotypes.py:336-358 and
:267-299).  VCF import is a plain-text reader (no pysam in this image).

On top of that, `hot_path_index()` flattens the `var2varid` dict once into the arrays the CUDA row builder
and the probability-table kernel consume (64-bit sorted keys, variant -> SNP ids, SNP -> variants CSR).
"""
from __future__ import annotations

import gzip
from collections import Counter, defaultdict
from copy import deepcopy
from typing import Dict, List, Tuple
from warnings import warn

import numpy as np
import pandas as pd

from .calls import BASE_TO_INDEX

_INITIAL_CAPACITY = 32768  # genotypes.py:33


class ProbabilisticGenotypes:
    def __init__(self, genotype_names: List[str], default_prior: float = 1.):
        self.var2varid: Dict[Tuple, int] = {}  # (chrom, pos0, base) -> row of variant_betas
        self.genotype_names: List[str] = list(genotype_names)
        assert self.genotype_names == sorted(self.genotype_names), 'please order genotype names'
        assert len(set(self.genotype_names)) == len(self.genotype_names), \
            f'Duplicates in genotypes: {genotype_names}'
        self.variant_betas: np.ndarray = np.zeros([_INITIAL_CAPACITY, self.n_genotypes], dtype='float32')
        self.default_prior: float = default_prior

    def __repr__(self):
        contigs = {chrom for chrom, _pos, _base in self.var2varid}
        return (f'<Genotypes with {self.n_variants} variants on {len(contigs)} contigs ("chromosomes") '
                f'and {self.n_genotypes} genotypes: \n{self.genotype_names}')

    # ------------------------------------------------------------------ basic accessors
    @property
    def n_genotypes(self) -> int:
        return len(self.genotype_names)

    @property
    def n_variants(self) -> int:
        return len(self.var2varid)

    def get_betas(self) -> np.ndarray:
        """Read-only [n_variants, n_genotypes] float32 view (genotypes.py:50-54)."""
        view = self.variant_betas[:self.n_variants]
        view.flags.writeable = False
        return view

    def get_snp_ids_for_variants(self) -> np.ndarray:
        """variant row -> dense id of its genomic position, first-seen order (genotypes.py:56-66)."""
        return self.hot_path_index()['variant2snp'].copy()

    def extend_variants(self, n_samples: int = 1) -> None:
        while self.n_variants + n_samples > len(self.variant_betas):
            self.variant_betas = np.concatenate([self.variant_betas, np.zeros_like(self.variant_betas)], axis=0)

    def get_variant_id(self, chrom, pos, base) -> int:
        key = (chrom, pos, base)
        vid = self.var2varid.get(key)
        if vid is None:
            vid = self.n_variants
            self.extend_variants(1)
            self.var2varid[key] = vid
        return vid

    def get_chromosome2positions(self) -> Dict[str, np.ndarray]:
        per_chrom = defaultdict(list)
        for chrom, pos, _base in self.var2varid:
            per_chrom[chrom].append(pos)
        if not per_chrom:
            warn('Genotypes are empty. Did you forget to add vcf/betas?')
        return {chrom: np.unique(np.asarray(positions, dtype=int)) for chrom, positions in per_chrom.items()}

    def get_snp_positions_set(self) -> set:
        return {(chrom, pos) for chrom, pos, _base in self.var2varid}

    def clone(self) -> 'ProbabilisticGenotypes':
        return deepcopy(self)

    def __deepcopy__(self, memo):
        out = ProbabilisticGenotypes.__new__(ProbabilisticGenotypes)
        memo[id(self)] = out
        for name, value in self.__dict__.items():
            if name == '_hot_index_cache':
                continue  # carried over below
            if name == 'var2varid':
                out.var2varid = dict(value)  # keys are immutable (chrom, pos, base) tuples: no need to recurse
                continue
            setattr(out, name, deepcopy(value, memo))
        cached = self.__dict__.get('_hot_index_cache')
        if cached is not None and cached['stamp'] == (id(self.var2varid), len(self.var2varid)):
            # same variants, same index (its arrays are never written); it is dropped as soon as the copy grows
            out.__dict__['_hot_index_cache'] = dict(cached, stamp=(id(out.var2varid), len(out.var2varid)))
        return out

    def _with_betas(self, external_betas: np.ndarray) -> 'ProbabilisticGenotypes':
        """Copy of the genotypes with replaced weights (genotypes.py:327-334)."""
        assert external_betas.shape == (self.n_variants, self.n_genotypes)
        assert external_betas.dtype == self.variant_betas.dtype
        assert np.min(external_betas) >= 0
        out = self.clone()
        out.variant_betas = external_betas.copy()
        return out

    # ------------------------------------------------------------------ flattened index for the CUDA path
    def hot_path_index(self) -> dict:
        """
        One pass over `var2varid` producing everything the device kernels need:
          keys_sorted  int64 [V]  (chrom_id << 40 | pos << 8 | base_code), ascending
          vids_sorted  int32 [V]  variant id of each sorted key
          chrom2id     dict       chromosome name -> chrom_id (first-seen order)
          variant2snp  int32 [V]  (demuxalot/genotypes.py:56-66)
          snp_offsets  int32 [S+1], snp_variants int32 [V]  -- CSR SNP -> its variants, ascending variant id
        Cached; the cache is invalidated when the dict object or its size changes.
        """
        stamp = (id(self.var2varid), len(self.var2varid))
        cached = self.__dict__.get('_hot_index_cache')
        if cached is not None and cached['stamp'] == stamp:
            return cached
        n = len(self.var2varid)
        # vectorised: the per-item Python loop cost 2.9 s at 5 M variants (factorize numbers in first-seen order,
        # exactly what the reference's setdefault(len(...)) loops do, genotypes.py:56-66)
        if n:
            chroms, positions, bases = zip(*self.var2varid.keys())
            vids = np.fromiter(self.var2varid.values(), dtype=np.int64, count=n)
            chrom_codes, chrom_names = pd.factorize(np.asarray(chroms, dtype=object), sort=False)
            pos = np.asarray(positions, dtype=np.int64)
            base_codes = np.fromiter((BASE_TO_INDEX[b] for b in bases), dtype=np.int64, count=n)
        else:
            vids = pos = base_codes = chrom_codes = np.zeros(0, dtype=np.int64)
            chrom_names = []
        chrom2id: Dict[object, int] = {name: k for k, name in enumerate(chrom_names)}
        keys = (chrom_codes.astype(np.int64) << 40) | ((pos & 0xFFFFFFFF) << 8) | base_codes
        snp_codes, snp_uniques = pd.factorize((chrom_codes.astype(np.int64) << 40) | (pos & 0xFFFFFFFFFF), sort=False)
        # demux.py:317-318 -- ids must enumerate the rows of variant_betas
        assert np.array_equal(np.sort(vids), np.arange(n)), 'variant ids must be a permutation of 0..V-1'
        variant2snp = np.full(n, -1, dtype=np.int32)
        variant2snp[vids] = snp_codes
        assert np.all(variant2snp >= 0)
        vids = vids.astype(np.int32)
        order = np.argsort(keys, kind='stable')
        n_snps = len(snp_uniques)
        snp_variants = np.argsort(variant2snp, kind='stable').astype(np.int32)
        snp_offsets = np.zeros(n_snps + 1, dtype=np.int32)
        np.cumsum(np.bincount(variant2snp, minlength=n_snps), out=snp_offsets[1:])
        cached = dict(stamp=stamp, keys_sorted=keys[order], vids_sorted=vids[order], chrom2id=chrom2id,
                      variant2snp=variant2snp, snp_offsets=snp_offsets, snp_variants=snp_variants, n_snps=n_snps)
        self.__dict__['_hot_index_cache'] = cached
        return cached

    # ------------------------------------------------------------------ importers
    def _check_imported_genotypes(self, imported_genotypes: List[str], allow_duplicates: bool = False) -> Dict[str, int]:
        repeated = [name for name, cnt in Counter(imported_genotypes).items() if cnt != 1]
        if repeated:
            if not allow_duplicates:
                raise RuntimeError(f'Duplicate genotypes found in imported data: {repeated}')
            warn(f'Duplicate genotypes found will be imported: {repeated}')
        imported, known = set(imported_genotypes), set(self.genotype_names)
        shared = imported & known
        if not shared:
            raise RuntimeError(f'No genotypes to import, expected {known}, got {imported}')
        if imported - known:
            warn(f'Genotypes will not be imported: {imported - known}')
        if known - imported:
            print(f'Some of genotypes are not provided during import: {known - imported}')
        return {name: self.genotype_names.index(name) for name in shared}

    def add_vcf(self, vcf_file_name, prior_strength: float = 100.) -> None:
        """
        Import diploid GT calls from a (optionally gzipped) text VCF; semantics of genotypes.py:112-168:
        every allele of a bi/multi-allelic SNV becomes a variant, each donor's prior_strength is split over
        its called alleles, donors without a call get 0.1 x the mean of the called donors, and records with
        fewer than two called donors are skipped.
        """
        opener = gzip.open if str(vcf_file_name).endswith('.gz') else open
        samples: List[str] = []
        donor2col = None
        n_records = n_skipped = 0
        n_before = self.n_variants
        with opener(vcf_file_name, 'rt') as stream:
            for line in stream:
                if line.startswith('##'):
                    continue
                fields = line.rstrip('\n').split('\t')
                if line.startswith('#CHROM'):
                    samples = fields[9:]
                    continue
                if len(fields) < 10:
                    continue
                n_records += 1
                chrom, pos1, ref, alt, fmt = fields[0], int(fields[1]), fields[3], fields[4], fields[8]
                alleles = [ref] + [a for a in alt.split(',') if a != '.']
                if any(len(a) != 1 for a in alleles):
                    print('skipping non-snp, alleles = ', tuple(alleles), chrom, pos1)
                    continue
                if donor2col is None:
                    donor2col = self._check_imported_genotypes(list(samples))
                if len(set(alleles)) != len(alleles) or any(a not in 'ACGT' for a in alleles):
                    n_skipped += 1
                    continue
                gt_slot = fmt.split(':').index('GT')
                rows = [self.get_variant_id(chrom, pos1 - 1, a) for a in alleles]
                contribution = np.zeros([len(rows), self.n_genotypes], dtype='float32')
                for donor, col in donor2col.items():
                    gt = fields[9 + samples.index(donor)].split(':')[gt_slot]
                    called = gt.replace('|', '/').split('/')
                    for allele_code in called:
                        if allele_code != '.':
                            contribution[int(allele_code), col] += prior_strength / len(called)
                missing = contribution.sum(axis=0) == 0
                if np.sum(~missing) < 2:
                    n_skipped += 1
                    continue
                contribution[:, missing] = contribution[:, ~missing].mean(axis=1, keepdims=True) * 0.1
                self.variant_betas[rows] += contribution
        if n_skipped:
            print('skipped', n_skipped, 'SNVs')
        print(f'Parsed {n_records} SNPs, got {self.n_variants - n_before} novel variants')

    def add_assignment_dataframe(self, assignment: pd.DataFrame, *, prior_stength: float = 100.) -> None:
        """
        columns = donors, index levels CHROM, POS1BASED, REF, ALT, values './.', '0/0', '0/1', '1/1' / None
        (genotypes.py:170-205).
        """
        assignment = assignment.fillna('./.')
        assignment.index = pd.MultiIndex.from_frame(
            assignment.index.to_frame().loc[:, ['CHROM', 'POS1BASED', 'REF', 'ALT']])
        donor2col = self._check_imported_genotypes(list(assignment.columns))
        assignment = assignment.loc[:, list(donor2col)]
        n_before = self.n_variants
        split = {'0/0': (1., 0.), '0/1': (.5, .5), '1/1': (0., 1.)}
        for (chrom, pos1, ref, alt), calls in assignment.iterrows():
            ref_row = self.get_variant_id(chrom, pos1 - 1, ref)
            alt_row = self.get_variant_id(chrom, pos1 - 1, alt)
            for donor, value in calls.items():
                if value in split:
                    to_ref, to_alt = split[value]
                    if to_ref:
                        self.variant_betas[ref_row, donor2col[donor]] += prior_stength * to_ref
                    if to_alt:
                        self.variant_betas[alt_row, donor2col[donor]] += prior_stength * to_alt
                else:
                    assert value == './.' or value is None, \
                        f'Unknown value: {value} of type {type(value)} at {chrom} {pos1} {ref} {alt}'
        print(f'Parsed {len(assignment) * 2} variants, of them  {self.n_variants - n_before} are novel')

    def add_prior_betas(self, prior_filename, *, prior_strength: float = 1.) -> None:
        """Accumulate a betas parquet written by `save_betas` (genotypes.py:267-299)."""
        frame: pd.DataFrame = pd.read_parquet(prior_filename) * prior_strength
        print('Provided prior information about genotypes:', [*frame.columns])
        absent = [name for name in self.genotype_names if name not in frame.columns]
        if absent:
            print(f'No information for genotypes: {absent}')
        index = frame.index.to_frame()
        rows = np.empty(len(frame), dtype=np.int64)
        for k, key in enumerate(zip(index['CHROM'], index['POS'], index['BASE'])):
            rows[k] = self.get_variant_id(*key)
        for col, donor in enumerate(self.genotype_names):
            if donor in frame.columns:
                np.add.at(self.variant_betas[:, col], rows, frame[donor].to_numpy())

    # ------------------------------------------------------------------ exporters
    def as_pandas_dataframe(self) -> pd.DataFrame:
        """Rows in tuple order of (chrom, pos, base); MultiIndex CHROM/POS/BASE (genotypes.py:336-354)."""
        ordered = sorted(self.var2varid.items())
        rows = np.asarray([vid for _key, vid in ordered], dtype=np.int64)
        index = pd.DataFrame({
            'CHROM': [key[0] for key, _ in ordered],
            'POS': [key[1] for key, _ in ordered],
            'BASE': [key[2] for key, _ in ordered],
        })
        betas = self.variant_betas[:self.n_variants][rows] if len(rows) else self.variant_betas[:0]
        return pd.DataFrame(data=betas, index=pd.MultiIndex.from_frame(index), columns=self.genotype_names)

    def save_betas(self, path_or_buf) -> None:
        """Learnt genotypes as a betas parquet; can be fed back with `add_prior_betas` (genotypes.py:356-358)."""
        self.as_pandas_dataframe().to_parquet(path_or_buf)
