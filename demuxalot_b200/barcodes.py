"""
Barcode whitelist -> dense integer ids.

Host-side mirror of the reference's `BarcodeHandler` (demuxalot/utils.py:39-109).  The hot path only
needs `n_barcodes` and `ordered_barcodes`; `get_barcode_index` is kept so
an input stage written against the reference keeps working (duck-typed reads with has_tag/get_tag).
"""
from __future__ import annotations

from collections import Counter
from pathlib import Path
from typing import Optional


class BarcodeHandler:
    def __init__(self, barcodes, RG_tags=None, tag: str = 'CB'):
        if isinstance(barcodes, (str, Path)):
            raise AssertionError('construct by passing list of possible barcodes')
        entries = list(barcodes)
        self.use_rg = RG_tags is not None
        if self.use_rg:
            rgs = list(RG_tags)
            assert len(rgs) == len(entries), 'RG tags should be the same length as barcodes'
            entries = list(zip(entries, rgs))
        assert len(set(entries)) == len(entries), 'all passed barcodes should be unique'
        # ids are positions in the sorted whitelist (utils.py:60-61)
        self.ordered_barcodes = sorted(entries)
        self.barcode2index = {bc: k for k, bc in enumerate(self.ordered_barcodes)}
        self.tag = tag

    @property
    def n_barcodes(self) -> int:
        return len(self.barcode2index)

    def get_barcode_index(self, read) -> Optional[int]:
        """None when the read has no barcode tag or the barcode is not whitelisted (utils.py:68-77)."""
        if not read.has_tag(self.tag):
            return None
        key = read.get_tag(self.tag)
        if self.use_rg:
            key = (key, read.get_tag('RG'))
        return self.barcode2index.get(key)

    @staticmethod
    def from_file(barcodes_filename, **kwargs) -> 'BarcodeHandler':
        """One barcode per line, optionally gzipped (utils.py:79-86)."""
        import pandas as pd
        column = pd.read_csv(barcodes_filename, header=None)[0]
        return BarcodeHandler(column.values.astype('str'), **kwargs)

    def filter_to_rg_value(self, rg_value) -> 'BarcodeHandler':
        """Handler for one RG of a merged BAM; other entries keep their slot under a dummy key (utils.py:88-99)."""
        assert self.use_rg
        sub = BarcodeHandler.__new__(BarcodeHandler)
        sub.tag = self.tag
        sub.use_rg = False
        sub.barcode2index = {
            (bc if rg == rg_value else idx): idx for (bc, rg), idx in self.barcode2index.items()
        }
        sub.ordered_barcodes = list(sub.barcode2index)
        return sub

    def __repr__(self):
        if not self.use_rg:
            return f'<BarcodeHandler with {self.n_barcodes} barcodes>'
        per_rg = Counter(rg for _bc, rg in self.barcode2index)
        return f'<BarcodeHandler with {self.n_barcodes} barcodes. Number of barcodes for RG codes: {per_rg}>'
