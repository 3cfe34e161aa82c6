"""
`CompressedSNPCalls` -- the hot path's input container (one per chromosome).

Mirrors the layout of the reference's class (demuxalot/snp_counter.py:77-139) so that objects produced by
the reference's `count_snps` and objects produced here are interchangeable:

  molecules  structured (compressed_cb i4, compressed_ub i4, p_group_misaligned f4)        12 B / record
  snp_calls  structured (molecule_index i4, snp_position i4, base_index u1, p_base_wrong f4) 13 B / record, packed

Both arrays may be over-allocated; only the first n_molecules / n_snp_calls entries are valid.
The device row builder consumes the raw packed bytes of these arrays directly (csrc/builder.cu).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np

MOLECULE_DTYPE = np.dtype([('compressed_cb', 'int32'), ('compressed_ub', 'int32'), ('p_group_misaligned', 'float32')])
SNP_CALL_DTYPE = np.dtype([('molecule_index', 'int32'), ('snp_position', 'int32'), ('base_index', 'uint8'),
                           ('p_base_wrong', 'float32')])
assert MOLECULE_DTYPE.itemsize == 12 and SNP_CALL_DTYPE.itemsize == 13

BASE_TO_INDEX = {'A': 0, 'C': 1, 'G': 2, 'T': 3, 'N': 4}  # demuxalot/utils.py:24


def _blank(dtype: np.dtype, n: int, fill: tuple) -> np.ndarray:
    out = np.empty(n, dtype=dtype)
    out[:] = fill
    return out


class CompressedSNPCalls:
    def __init__(self, start_snps_size: int = 1024, start_molecule_size: int = 128):
        self.n_molecules = 0
        self.molecules = _blank(MOLECULE_DTYPE, start_molecule_size, (-1, -1, -1.))
        self.n_snp_calls = 0
        self.snp_calls = _blank(SNP_CALL_DTYPE, start_snps_size, (-1, -1, 255, -1.))

    # -- incremental filling, as the reference's input stage does (snp_counter.py:100-112) ---------------
    def add_calls_from_read_group(self, compressed_cb: int, compressed_ub: int, p_group_misaligned: float,
                                  snps: Sequence[Tuple[int, str, float]]) -> None:
        while self.n_snp_calls + len(snps) > len(self.snp_calls):
            self.snp_calls = np.concatenate([self.snp_calls, self.snp_calls])
        if self.n_molecules == len(self.molecules):
            self.molecules = np.concatenate([self.molecules, self.molecules])
        mol = self.n_molecules
        self.molecules[mol] = (compressed_cb, compressed_ub, p_group_misaligned)
        self.n_molecules += 1
        for position, base, p_wrong in snps:
            self.snp_calls[self.n_snp_calls] = (mol, position, BASE_TO_INDEX[base], p_wrong)
            self.n_snp_calls += 1

    # -- bulk construction (synthetic generators, tests) ---------------------------------------------------
    @classmethod
    def from_arrays(cls, compressed_cb, compressed_ub, p_group_misaligned,
                    molecule_index, snp_position, base_index, p_base_wrong,
                    spare_capacity: int = 0) -> 'CompressedSNPCalls':
        out = cls.__new__(cls)
        n_mol, n_calls = len(compressed_cb), len(molecule_index)
        out.molecules = _blank(MOLECULE_DTYPE, n_mol + spare_capacity, (-1, -1, -1.))
        out.molecules['compressed_cb'][:n_mol] = compressed_cb
        out.molecules['compressed_ub'][:n_mol] = compressed_ub
        out.molecules['p_group_misaligned'][:n_mol] = p_group_misaligned
        out.snp_calls = _blank(SNP_CALL_DTYPE, n_calls + spare_capacity, (-1, -1, 255, -1.))
        out.snp_calls['molecule_index'][:n_calls] = molecule_index
        out.snp_calls['snp_position'][:n_calls] = snp_position
        out.snp_calls['base_index'][:n_calls] = base_index
        out.snp_calls['p_base_wrong'][:n_calls] = p_base_wrong
        out.n_molecules, out.n_snp_calls = n_mol, n_calls
        return out

    def minimize_memory_footprint(self) -> None:
        self.snp_calls = self.snp_calls[:self.n_snp_calls].copy()
        self.molecules = self.molecules[:self.n_molecules].copy()
        assert np.all(self.molecules['p_group_misaligned'] != -1)
        assert np.all(self.snp_calls['p_base_wrong'] != -1)

    @staticmethod
    def concatenate(snp_calls_list: Iterable['CompressedSNPCalls']) -> 'CompressedSNPCalls':
        """Merge containers of the same chromosome; molecule indices are rebased (snp_counter.py:120-139)."""
        parts = list(snp_calls_list)
        out = CompressedSNPCalls.__new__(CompressedSNPCalls)
        if not parts:
            out.molecules = _blank(MOLECULE_DTYPE, 0, (-1, -1, -1.))
            out.snp_calls = _blank(SNP_CALL_DTYPE, 0, (-1, -1, 255, -1.))
        else:
            # every record is copied once, straight into its place in the result
            out.molecules = np.empty(sum(p.n_molecules for p in parts), dtype=parts[0].molecules.dtype)
            out.snp_calls = np.empty(sum(p.n_snp_calls for p in parts), dtype=parts[0].snp_calls.dtype)
            base = first_call = 0
            for part in parts:
                out.molecules[base:base + part.n_molecules] = part.molecules[:part.n_molecules]
                block = out.snp_calls[first_call:first_call + part.n_snp_calls]
                block[...] = part.snp_calls[:part.n_snp_calls]
                if base:
                    block['molecule_index'] += base
                base += part.n_molecules
                first_call += part.n_snp_calls
        out.n_molecules, out.n_snp_calls = len(out.molecules), len(out.snp_calls)
        return out
