import sys, time
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import numpy as np
import pysam_shim
sys.modules['pysam']=pysam_shim
sys.path.insert(0,'/root/reference')
import demuxalot as ref
from demuxalot_b200 import BarcodeHandler, ProbabilisticGenotypes
from demuxalot_b200.counting import count_snps
D='/root/reference/examples/example_data/'
g=ProbabilisticGenotypes(['Donor01','Donor02','Donor03','Donor04']); g.add_vcf(D+'test_genotypes.vcf')
bh=BarcodeHandler.from_file(D+'test_barcodes.csv')
t=time.time(); mine=count_snps(D+'test_bamfile.bam', g.get_chromosome2positions(), bh, joblib_n_jobs=1); t1=time.time()-t
rg=ref.ProbabilisticGenotypes(['Donor01','Donor02','Donor03','Donor04']); rg.add_vcf(D+'test_genotypes.vcf')
rbh=ref.BarcodeHandler.from_file(D+'test_barcodes.csv')
t=time.time(); theirs=ref.count_snps(D+'test_bamfile.bam', rg.get_chromosome2positions(), rbh, joblib_n_jobs=1, joblib_verbosity=0); t2=time.time()-t
print('mine %.1fs reference(shim) %.1fs'%(t1,t2), list(mine), list(theirs))
ok = list(mine)==list(theirs)
for c in theirs:
    a,b=mine[c],theirs[c]
    ok &= a.n_molecules==b.n_molecules and a.n_snp_calls==b.n_snp_calls
    ok &= np.array_equal(a.molecules[:a.n_molecules], b.molecules[:b.n_molecules])
    ok &= np.array_equal(a.snp_calls[:a.n_snp_calls], b.snp_calls[:b.n_snp_calls])
    print(c, a.n_molecules, a.n_snp_calls, ok)
print('IDENTICAL' if ok else 'DIFFERENT')
