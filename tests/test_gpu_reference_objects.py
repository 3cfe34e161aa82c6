"""
The CUDA path fed with the reference's OWN objects (missing #9): tests/golden/ref_objects.pkl holds
demuxalot.ProbabilisticGenotypes / BarcodeHandler / CompressedSNPCalls instances pickled by the unmodified reference
(tests/golden/make_ref_objects.py) and the reference's outputs for them.  The GPU box has no reference package, so
stand-in classes with the reference's accessors (cited) are registered under its module paths before unpickling; the
restored objects carry exactly the reference's state and go through `Demultiplexer` by duck typing
(`_foreign_hot_index`, demuxalot_b200/demultiplexer.py).
"""
import pickle
import sys
import types
from copy import deepcopy
from pathlib import Path

import numpy as np
import pytest

FIXTURE = Path(__file__).resolve().parent / 'golden' / 'ref_objects.pkl'


def _load_payload():
    if 'demuxalot' in sys.modules and hasattr(sys.modules['demuxalot'], 'Demultiplexer'):
        with open(FIXTURE, 'rb') as f:  # the real reference is importable (build container): use it
            return pickle.load(f)

    class ProbabilisticGenotypes:  # accessors of demuxalot/genotypes.py:38-54, 327-334, 360-361
        n_genotypes = property(lambda self: len(self.genotype_names))
        n_variants = property(lambda self: len(self.var2varid))

        def get_betas(self):
            view = self.variant_betas[:self.n_variants]
            view.flags.writeable = False
            return view

        def _with_betas(self, external_betas):
            assert external_betas.shape == (self.n_variants, self.n_genotypes)
            assert external_betas.dtype == self.variant_betas.dtype
            assert np.min(external_betas) >= 0
            result = deepcopy(self)
            result.variant_betas = external_betas.copy()
            return result

    class BarcodeHandler:  # demuxalot/utils.py:64-66
        n_barcodes = property(lambda self: len(self.barcode2index))

    class CompressedSNPCalls:  # demuxalot/snp_counter.py:77-98: plain state
        pass

    saved = {name: sys.modules.get(name) for name in ('demuxalot', 'demuxalot.genotypes', 'demuxalot.utils',
                                                      'demuxalot.snp_counter')}
    try:
        for name, cls in (('demuxalot.genotypes', ProbabilisticGenotypes), ('demuxalot.utils', BarcodeHandler),
                          ('demuxalot.snp_counter', CompressedSNPCalls)):
            module = types.ModuleType(name)
            setattr(module, cls.__name__, cls)
            cls.__module__ = name
            sys.modules[name] = module
        sys.modules['demuxalot'] = types.ModuleType('demuxalot')
        with open(FIXTURE, 'rb') as f:
            return pickle.load(f)
    finally:
        for name, module in saved.items():
            if module is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = module


def test_fixture_restores_the_reference_state():
    payload = _load_payload()
    g, h = payload['genotypes'], payload['barcode_handler']
    assert type(g).__module__ == 'demuxalot.genotypes' and type(h).__module__ == 'demuxalot.utils'
    assert g.n_variants == len(g.var2varid) and g.variant_betas.dtype == np.float32 and g.default_prior == 0.7
    assert h.n_barcodes == len(h.ordered_barcodes) == payload['expected']['logits'].shape[0]
    for calls in payload['calls'].values():
        assert type(calls).__module__ == 'demuxalot.snp_counter'
        assert calls.snp_calls.dtype.itemsize == 13 and calls.molecules.dtype.itemsize == 12


@pytest.mark.gpu
def test_reference_objects_through_the_cuda_path(native_lib):
    from demuxalot_b200 import Demultiplexer as D
    payload = _load_payload()
    g, h, calls, want = payload['genotypes'], payload['barcode_handler'], payload['calls'], payload['expected']
    assert not hasattr(g, 'hot_path_index')  # a foreign object: the index is derived from var2varid
    logits, post = D.predict_posteriors(calls, g, h, doublet_prior=0.35)
    assert list(logits.columns) == want['columns'] and list(logits.index) == want['index']
    assert logits.index.name == 'BARCODE'
    rel = np.abs(logits.values.astype(np.float64) - want['logits']) / np.maximum(np.abs(want['logits']), 1e-30)
    assert rel.max() <= 1e-5
    dlogit = np.abs(logits.values.astype(np.float64) - want['logits']).max(axis=1, keepdims=True)
    assert (np.abs(post.values - want['posteriors']) <= 1e-6 + 0.5 * dlogit).all()
    learnt, learnt_post = D.learn_genotypes(calls, g, h, n_iterations=4, doublet_prior=0.35)
    assert type(learnt) is type(g)  # the reference's own _with_betas built the result
    got = np.array(learnt.get_betas(), np.float64)
    assert (np.abs(got - want['learnt_betas']) / np.maximum(np.abs(want['learnt_betas']), 1e-3)).max() <= 1e-5
    assert np.abs(learnt_post.values - want['learnt_posteriors']).max() <= 1e-5
